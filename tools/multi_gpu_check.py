"""N-GPU check of the sharded forward (run under torchrun, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Every rank computes the whole global batch locally as the reference and compares the gathered tensor of both gather
implementations (copy-engine peer reads / NCCL all_gather) bit for bit, over several steps (ring reuse) and through the
host-buffer serving call.  Test tooling."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import migan_b200  # noqa: E402
from migan_b200 import parallel, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
    parallel.configure_overlap(gather=mode)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    R, n = 128, 4
    g = migan_b200.Generator(R, path="tc")
    g.load_state_dict(synthetic.export_style_state_dict(R, seed=1))
    g = g.to(dev).eval()
    g.graph_max_batch = 0
    sg = parallel.ShardedGenerator(g, gather=mode)
    sg.check_replicas(g.state_dict())
    ok = True
    handles = []
    refs = []
    for step in range(7):
        xg = synthetic.synthetic_input(R, n * world, seed=100 + step).to(dev)     # same global batch on every rank
        refs.append(g(xg))
        handles.append(sg.forward_async(xg[rank * n:(rank + 1) * n].contiguous()))
        if len(handles) > 2:                                                       # two gathers in flight, like bench.py
            ok = ok and torch.equal(handles.pop(0).wait(), refs.pop(0))
    while handles:
        ok = ok and torch.equal(handles.pop(0).wait(), refs.pop(0))
    # host-buffer serving path
    xs = [synthetic.synthetic_input(R, n * world, seed=200 + s) for s in range(4)]
    outs = [torch.empty(n, 3, R, R).pin_memory() for _ in range(4)]
    for xg, o in zip(xs, outs):
        sg.forward_host_async(xg[rank * n:(rank + 1) * n].contiguous().pin_memory(), o)
    sg.host_wait()
    for xg, o in zip(xs, outs):
        ok = ok and torch.equal(o, g(xg[rank * n:(rank + 1) * n].contiguous().to(dev)).cpu())
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi_gpu_check gather=%s (effective %s) world=%d: %s" % (mode, sg.gather if sg._ce is None else "ce", world, "OK" if t.item() == 1.0 else "MISMATCH"), flush=True)
    sg.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
