"""Worker of tests/test_gpu_dp.py (launched by torch.distributed.run, one process per GPU).

Checks the data-parallel gradient sync of host/convnet.cc (DataParallelSync::AllReduceAverageAsync + WaitAll, the NCCL
replacement of the reference's ConvNet::Accumulate/Broadcast, src/convnet.cc:407-450) ON HARDWARE:
  1. after 3 training steps the parameters are BIT-IDENTICAL on every rank (the property the reference relies on,
     convnet.cc:442-447);
  2. they equal (<= 1e-5 relative to the largest parameter) a 1-rank run of the same model on the concatenated batch —
     the all-reduce averages per-rank means, i.e. the global-batch mean;
  3. the bucket / event ordering is exercised with several bucket sizes (one bucket, one bucket per edge).
fp32 conv arithmetic, nets without dropout ("tiny": conv, max-pool, rnorm, 1x1, strided conv, avg-pool, fc).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from convnet_b200 import lib  # noqa: E402
from convnet_b200.net import Net, dp_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    lib.set_precision(os.environ.get("DP_PRECISION", "fp32"))
    model, B, steps = os.environ.get("DP_MODEL", "tiny"), int(os.environ.get("DP_BATCH", "32")), 3
    results = []
    for bucket_floats in (1 << 30, 1, 4096):                     # one bucket / one per edge / mixed
        net = Net(model, B, seed=7)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dp_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        net.dp_init(rank, world, bytes(idt.cpu().numpy().tobytes()), bucket_floats)
        F = net.input_floats // B
        g = torch.Generator(device="cuda").manual_seed(99)
        xg = torch.randn(steps, F, world * B, device="cuda", generator=g)          # [step][feature][global image]
        yg = torch.randint(0, net.num_classes, (steps, world * B), device="cuda", generator=g, dtype=torch.int32)
        for s in range(steps):
            net.input_tensor().copy_(xg[s][:, rank * B:(rank + 1) * B].contiguous().view(-1))
            net.labels_tensor().copy_(yg[s][rank * B:(rank + 1) * B])
            net.train_step(False)
        torch.cuda.synchronize()
        p = net.params_tensor().clone()
        gathered = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(gathered, p)
        identical = all(torch.equal(gathered[0], t) for t in gathered)
        rel = None
        if rank == 0:
            ref = Net(model, world * B, seed=7)                                     # same seed -> same initial parameters
            for s in range(steps):
                ref.input_tensor().copy_(xg[s].contiguous().view(-1))
                ref.labels_tensor().copy_(yg[s])
                ref.train_step(False)
            torch.cuda.synchronize()
            q = ref.params_tensor()
            rel = ((p - q).abs().max() / q.abs().max()).item()
            moved = ((q - Net(model, world * B, seed=7).params_tensor()).abs().max()).item()   # training changed something
            ref.close()
            results.append({"bucket_floats": bucket_floats, "bit_identical_across_ranks": identical,
                            "rel_diff_vs_1rank_global_batch": rel, "max_param_change": moved})
        ok = torch.tensor([1 if identical and (rel is None or rel < 1e-5) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        net.close()
        if ok.item() != 1:
            if rank == 0:
                print(json.dumps({"ok": False, "results": results}), flush=True)
            dist.destroy_process_group()
            sys.exit(1)
    if rank == 0:
        print(json.dumps({"ok": True, "world": world, "model": model, "per_rank_batch": B, "results": results}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
