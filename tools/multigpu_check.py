"""Multi-GPU equivalence check (run under torchrun, N = 2, 4 or 8 GPUs of one box):
CFG split x frame sharding must reproduce the single-GPU noise prediction of the same inputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [H W]
"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavie_b200 import UNet3DConditionModel
from lavie_b200.synthetic import synthetic_inputs, synthetic_state_dict


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    graph = (sys.argv[3] != "eager") if len(sys.argv) > 3 else True
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P = world // 2
    c, s = rank // P, rank % P
    frame_groups = [dist.new_group(list(range(cc * P, cc * P + P))) for cc in range(2)]
    unet = UNet3DConditionModel(use_cuda_graph=graph)
    unet.load_state_dict(synthetic_state_dict(seed=0), strict=True)
    unet = unet.to(dev).eval()
    F = 16
    sample, t, text = synthetic_inputs(2, F, H, W, seed=0)
    sample, text = sample.to(dev), text.to(dev)
    ref = unet(sample, t, encoder_hidden_states=text).sample                       # single-GPU answer, all frames
    backend = sys.argv[4] if len(sys.argv) > 4 else "p2p"
    unet.set_frame_sharding(frame_groups[c], backend=backend)
    fl = F // P
    shard = sample[c:c + 1, :, s * fl:(s + 1) * fl].contiguous()
    out = unet(shard, t, encoder_hidden_states=text[c:c + 1]).sample
    want = ref[c:c + 1, :, s * fl:(s + 1) * fl]
    err = float((out.double() - want.double()).norm() / want.double().norm())
    # timing of the sharded forward
    for _ in range(3):
        unet(shard, t, encoder_hidden_states=text[c:c + 1])
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        unet(shard, t, encoder_hidden_states=text[c:c + 1])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    errs = [None] * world
    dist.all_gather_object(errs, (rank, c, s, err, ms))
    if rank == 0:
        for r in errs:
            print(f"rank {r[0]} (cfg half {r[1]}, frame shard {r[2]}/{P}): rel-L2 vs single GPU = {r[3]:.3e}, {r[4]:.2f} ms/forward")
        worst = max(r[3] for r in errs)
        print(f"RESULT world={world} P={P} graph={graph} backend={backend} worst_rel_l2={worst:.3e} {'OK' if worst < 2e-2 else 'FAIL'}")
    torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush()
    os._exit(0)     # process-group teardown hangs with captured NCCL graphs


if __name__ == "__main__":
    main()
