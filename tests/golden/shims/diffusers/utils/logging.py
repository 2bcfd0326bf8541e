import logging as _l


def get_logger(name):
    return _l.getLogger(name)
