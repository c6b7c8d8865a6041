"""2-rank probe (torchrun): CUDA IPC handle exchange, peer-store flag latency, halo-row push, NCCL allreduce latency.
Run on the GPU box:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ipc_probe.py
"""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    gl = dist.new_group(backend="gloo")
    lib = ctypes.CDLL(os.path.join(HERE, "libipc_probe.so"))
    lib.probe_pingpong.restype = ctypes.c_double
    lib.probe_exchange.restype = ctypes.c_double
    lib.probe_exchange.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_ulonglong]
    lib.probe_init.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p]
    h = ctypes.create_string_buffer(64)
    rc = lib.probe_init(rank, 1 << 22, h)
    out = {"rank": rank, "init": rc}
    handles = [None] * world
    dist.all_gather_object(handles, bytes(h.raw), group=gl)
    peer = (rank + 1) % world
    rc = lib.probe_open(handles[peer])
    out["open"] = rc
    dist.barrier(group=gl)
    if rc == 0:
        out["pingpong_oneway_us"] = lib.probe_pingpong(rank, 2000)
        dist.barrier(group=gl)
        seq = 0
        for n in (1, 8192, 65536):
            out["exchange_%d_doubles_us" % n] = lib.probe_exchange(n, 500, seq)
            seq += 500
            dist.barrier(group=gl)
    # NCCL small allreduce latency, back to back on one stream
    t = torch.zeros(2, dtype=torch.float64, device="cuda")
    for _ in range(20):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(500):
        dist.all_reduce(t)
    b.record()
    torch.cuda.synchronize()
    out["nccl_allreduce_16B_us"] = a.elapsed_time(b) * 1000 / 500
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
