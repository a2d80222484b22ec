"""Measures, under torchrun, what the node-table exchange can cost on this box:
  1. NCCL all_gather_into_tensor of one meta-path's table (N x 72 fp32 split over the ranks)
  2. the same exchange as copy-engine pulls from peer-mapped symmetric memory (no SMs involved)
  3. whether NVLS multicast is available for symmetric memory
Run:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/comm_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as td

N_NODES = int(os.environ.get("PROBE_NODES", 2_000_000))
COLS = 72


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    td.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    td.init_process_group("nccl", device_id=dev)
    n_pad = -(-N_NODES // world)
    shard = torch.randn(n_pad, COLS, device=dev)
    full = torch.empty(world * n_pad, COLS, device=dev)
    recv_mb = (world - 1) * shard.numel() * 4 / 1e6

    ms = timed(lambda: td.all_gather_into_tensor(full, shard))
    if rank == 0:
        print(f"[probe] world={world} shard={shard.numel() * 4 / 1e6:.1f} MB  NCCL all_gather: {ms:.3f} ms "
              f"-> {recv_mb / ms:.1f} GB/s received per rank", flush=True)

    try:
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(n_pad, COLS, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, td.group.WORLD)
        buf.copy_(shard)
        mc = getattr(hdl, "multicast_ptr", 0)
        if rank == 0:
            print(f"[probe] symmetric memory ok; multicast support={getattr(hdl, 'has_multicast_support', None)} "
                  f"multicast_ptr={'nonzero' if mc else 0}", flush=True)
        peers = [hdl.get_buffer(r, (n_pad, COLS), torch.float32) for r in range(world)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(world)]

        def pull():
            cur = torch.cuda.current_stream()
            hdl.barrier(channel=0)
            for k in range(1, world):
                r = (rank + k) % world
                st = streams[k]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    full[r * n_pad:(r + 1) * n_pad].copy_(peers[r], non_blocking=True)
            for k in range(1, world):
                cur.wait_stream(streams[k])
            hdl.barrier(channel=1)

        ms = timed(pull)
        ok = torch.equal(full[((rank + 1) % world) * n_pad:((rank + 1) % world + 1) * n_pad], peers[(rank + 1) % world])
        if rank == 0:
            print(f"[probe] copy-engine pull from {world - 1} peers: {ms:.3f} ms -> {recv_mb / ms:.1f} GB/s "
                  f"received per rank (data ok={ok})", flush=True)
        # SM-based pull: an elementwise kernel that loads straight from the peer mapping (LDG over NVLink)
        def sm_pull():
            hdl.barrier(channel=0)
            for k in range(1, world):
                r = (rank + k) % world
                torch.mul(peers[r], 1.0, out=full[r * n_pad:(r + 1) * n_pad])
            hdl.barrier(channel=1)

        ms = timed(sm_pull)
        if rank == 0:
            print(f"[probe] SM pull (LDG from peers): {ms:.3f} ms -> {recv_mb / ms:.1f} GB/s received per rank", flush=True)

        # multicast push: every rank writes its shard once to the NVLS multicast address of a full table
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import ctypes
        from han_b200 import _lib
        fullsym = symm.empty(world * n_pad, COLS, dtype=torch.float32, device=dev)
        hf = symm.rendezvous(fullsym, td.group.WORLD)
        mcp = int(hf.multicast_ptr)
        if mcp:
            def mc_push():
                hf.barrier(channel=0)
                _lib.call("han_multicast_copy", _lib.ptr(shard), ctypes.c_void_p(mcp + rank * n_pad * COLS * 4),
                          n_pad * COLS, _lib.stream_ptr())
                hf.barrier(channel=1)

            ms = timed(mc_push)
            nxt = (rank + 1) % world
            td.all_gather_into_tensor(full, shard)
            torch.cuda.synchronize()
            ok = torch.equal(fullsym, full)
            if rank == 0:
                print(f"[probe] multicast push (multimem.st, coalesced, 37 CTAs): {ms:.3f} ms -> {recv_mb / ms:.1f} GB/s "
                      f"received per rank (data ok={ok})", flush=True)
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            import traceback
            print(f"[probe] symmetric memory unavailable: {type(ex).__name__}: {ex}\n{traceback.format_exc()}", flush=True)
    td.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
