"""Two-GPU probe of the transfer primitives the copy-engine all-gather is built from (run under torchrun, 2+ processes):
peer-to-peer copy bandwidth through CUDA-IPC mappings in both directions (torch copy_ vs a bare cudaMemcpyAsync), idle and
under a running forward, and the latency of the stream-memory-operation signal.  Measurement tooling."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import migan_b200  # noqa: E402
from migan_b200 import _abi  # noqa: E402


def timed(fn, stream, reps=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    from torch.multiprocessing.reductions import reduce_tensor
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load()
    MB = 100
    buf = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
    src = torch.ones(MB << 20, dtype=torch.uint8, device=dev)
    flag = torch.zeros(4, dtype=torch.int32, device=dev)
    pub = [None] * world
    dist.all_gather_object(pub, (reduce_tensor(buf), reduce_tensor(flag)))
    peer = (rank + 1) % world
    pbuf = pub[peer][0][0](*pub[peer][0][1])
    torch.cuda.synchronize(); dist.barrier()
    st = torch.cuda.Stream(device=dev)
    res = {}
    nb = MB << 20
    with torch.cuda.stream(st):
        res["pull torch copy_"] = timed(lambda: buf.copy_(pbuf, non_blocking=True), st)
        res["push torch copy_"] = timed(lambda: pbuf.copy_(src, non_blocking=True), st)
        res["pull cudaMemcpyAsync"] = timed(lambda: lib.b200_memcpy_async(buf.data_ptr(), pbuf.data_ptr(), nb, st.cuda_stream), st)
        res["push cudaMemcpyAsync"] = timed(lambda: lib.b200_memcpy_async(pbuf.data_ptr(), src.data_ptr(), nb, st.cuda_stream), st)
        res["local d2d"] = timed(lambda: lib.b200_memcpy_async(buf.data_ptr(), src.data_ptr(), nb, st.cuda_stream), st)
    dist.barrier()
    # under load: a forward running on the default stream of BOTH GPUs while the copies go
    g = migan_b200.Generator(512).to(dev).eval()
    x = torch.randn(32, 4, 512, 512, device=dev)
    for _ in range(2):
        g(x)
    torch.cuda.synchronize(); dist.barrier()
    for name, fn in (("push cudaMemcpyAsync under forward", lambda: lib.b200_memcpy_async(pbuf.data_ptr(), src.data_ptr(), nb, st.cuda_stream)),
                     ("pull cudaMemcpyAsync under forward", lambda: lib.b200_memcpy_async(buf.data_ptr(), pbuf.data_ptr(), nb, st.cuda_stream))):
        for _ in range(4):
            g(x)
        res[name] = timed(fn, st, reps=10)
        torch.cuda.synchronize(); dist.barrier()
    # forward time with and without a concurrent stream of pushes
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, with_copy in (("forward alone", False), ("forward + pushes", True)):
        torch.cuda.synchronize(); dist.barrier()
        a.record()
        for _ in range(5):
            g(x)
            if with_copy:
                lib.b200_memcpy_async(pbuf.data_ptr(), src.data_ptr(), nb, st.cuda_stream)
        b.record()
        torch.cuda.synchronize()
        res[label + " ms/step"] = a.elapsed_time(b) / 5
    # signal latency: stream A waits for flag[0] >= 1, stream B writes it
    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    lat = []
    for i in range(1, 6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.b200_stream_wait_value32(sa.cuda_stream, flag.data_ptr(), i)
        e1.record(sa)
        time.sleep(0.01)
        e0.record(sb)
        lib.b200_stream_write_value32(sb.cuda_stream, flag.data_ptr(), i)
        torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1) * 1e3)
    res["local write->wait latency us"] = sorted(lat)[len(lat) // 2]
    if rank == 0:
        for k, v in res.items():
            if "ms/step" in k or "latency" in k:
                print("%-44s %.3f" % (k, v), flush=True)
            else:
                print("%-44s %.3f ms  = %.0f GB/s" % (k, v, MB / 1024 / (v * 1e-3)), flush=True)
    del pbuf
    torch.cuda.synchronize(); dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
