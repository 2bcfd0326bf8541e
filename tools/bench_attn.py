import sys, torch
sys.path.insert(0, ".")
from lavie_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
heads = 8
for batch, Sq, Sk, d, pitch, div in [(32, 2560, 2560, 40, 48, 1), (32, 640, 640, 80, 80, 1), (32, 160, 160, 160, 160, 1),
                                     (32, 2560, 77, 40, 48, 16), (32, 640, 77, 80, 80, 16)]:
    hp = heads * pitch
    def mk(rows):
        t = torch.zeros(rows, heads, pitch, device=dev); t[..., :d] = torch.randn(rows, heads, d, device=dev)
        return t.reshape(rows, hp).to(torch.bfloat16)
    q, k, v = mk(batch * Sq), mk(batch // div * Sk), mk(batch // div * Sk)
    fl = 4.0 * batch * heads * Sq * Sk * d
    for poly in (0, 4):
        lib.lavie_debug_set(4, poly)
        ms = timeit(lambda: ops.attention(q, k, v, batch, heads, Sq, Sk, d, pitch, div))
        print(f"attn B={batch} Sq={Sq} Sk={Sk} d={d} poly={poly}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s")
    lib.lavie_debug_set(4, 0)
