#!/usr/bin/env python
"""Benchmark of the HAN hot path (BASELINE.json metric: fused HAN forward+backward meta-path edges/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload syn2m|acm|dblp|imdb|mag]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's dense algorithm on the host cores

A step = one forward + backward of HeteGAT_multi.inference + masked CE + L2 over the whole synthetic
graph (projection, node attention, semantic attention, classifier, loss, every gradient; optimizer
excluded).  edges = sum_p nnz(meta-path mask p), self-loops included, counted once per step.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "han_fwd_bwd_metapath_edges_per_s"
UNIT = "edges/s"
K_HEADS, HID, ATT = 8, 8, 128
PROJ_NAMES = {0: "fp32 FFMA", 1: "tcgen05 3xTF32 (fp32-grade)", 2: "tcgen05 2xTF32 (0/1 features exact; fp32-grade)",
              3: "tcgen05 TF32"}

WORKLOADS = {
    "syn2m": "synthetic 2M-node heterograph, 4 meta-paths, avg degree 50, 256-d feats",
    "mag": "OGB-MAG-scale synthetic (736k papers, 2 power-law meta-paths)",
    "acm": "ACM3025-shaped synthetic graph (3025 papers, 1870-d feats, PAP+PLP)",
    "dblp": "DBLP four-area-shaped synthetic (4057 authors, 334-d feats, APA/APCPA/APTPA)",
    "imdb": "IMDB-shaped synthetic (4780 movies, MAM/MDM meta-paths, 3 classes)",
}


# exact edge counts of the large synthetic graphs (a pure function of the generator spec and seed; measured by the
# GPU arm, which asserts them): lets the CPU reference arm echo the same `config` without generating 10^9 edges
KNOWN_EDGES = {"syn2m": 399_995_112, "mag": 1_004_743_520}     # what the seeded generators produce (checked at run time)


def pick_partition(args, world, n_paths):
    if world > 1 and args.partition == "auto":
        # (meta-path x row-block) tiles whenever the rank count and the meta-path count divide one another: they move
        # 4.4x fewer bytes between ranks than destination-row shards, and the Z re-sharding rides on K-B's stores
        return "tile" if (world % n_paths == 0 or n_paths % world == 0) else "row"
    return args.partition


def workload_config(args, world, edges=None):
    """The `config` object of the JSON line -- the same for both arms (the reference arm runs on OUR arm's config)."""
    from han_b200 import synth
    name = args.workload
    if name in synth.SMALL:
        cfg = synth.SMALL[name]()
        N, F, C, P = cfg.N, cfg.F, cfg.C, cfg.P
        edges = cfg.n_edges() if edges is None else edges
        small = True
    else:
        spec = synth.LARGE[name]
        N, F, C, P = spec.N, spec.F, spec.C, spec.P
        edges = KNOWN_EDGES.get(name) if edges is None else edges
        small = False
    pmode = projection_mode(args)
    part = pick_partition(args, world, P)
    return {"workload": WORKLOADS[name], "nodes": N, "features": F, "meta_paths": P, "edges": edges, "heads": K_HEADS,
            "hid": HID, "mp_att_size": ATT, "classes": C,
            "parallelism": (f"(meta-path x row-block) tiles x{world}" if part == "tile" else f"dst-row shards x{world}")
            if world > 1 else "single GPU",
            "l2_policy": "L2 flushed between timed steps" if small else "inputs larger than L2",
            "dropout": args.dropout,
            "projection": PROJ_NAMES[pmode] if not args.dropout else "fp32 FFMA with per-head input masks"}


def projection_mode(args):
    from han_b200 import synth
    return {"fp32": 0, "tf32x3": 1, "tf32x2": 2, "tf32": 3,
            "auto": 2 if args.workload in synth.SMALL else 1}[args.projection]   # SMALL configs: 0/1 features


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        return len(self.lines)

    def stop(self, lo=0, hi=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[lo:hi]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        hi = [s for s in sm if mx and s > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def build_workload(name, dev, dist):
    """-> dict(N, F, C, P, X (local rows, device), graphs (local dst rows), labels, mask, edges (global),
    host (pinned host copies for the e2e leg), working_set_bytes)."""
    import han_b200 as hb
    from han_b200 import synth
    lo, hi = (0, None)
    tile = dist if (dist is not None and hasattr(dist, "paths")) else None     # tiles.TileShard

    def spans(N, P):
        """-> (attention rows, label rows, meta-paths) of this rank."""
        if tile is not None:
            a, srows = tile.rows(N)
            return a, srows, tile.paths
        r = dist.row_range(N) if dist else (0, N)
        return r, r, list(range(P))

    if name in synth.SMALL:
        cfg = synth.SMALL[name]()
        N, F, C, P = cfg.N, cfg.F, cfg.C, cfg.P
        (lo, hi), (s_lo, s_hi), paths = spans(N, P)
        full = [hb.process.adj_to_bias(a, [N]) for a in cfg.adjs()]
        X = torch.from_numpy(cfg.X[lo:hi]).to(dev)
        labels = torch.from_numpy(cfg.labels[s_lo:s_hi]).to(dev)
        mask = torch.from_numpy(cfg.train_mask[s_lo:s_hi].astype(np.float32)).to(dev)
        edges = cfg.n_edges()
        host = {"X": torch.from_numpy(cfg.X[lo:hi]).pin_memory(),
                "bias": [torch.from_numpy(hb_bias(m[lo:hi])).pin_memory() for m in cfg.masks] if dist is None else None}
    else:
        spec = synth.LARGE[name]
        N, F, C, P = spec.N, spec.F, spec.C, spec.P
        (lo, hi), (s_lo, s_hi), paths = spans(N, P)
        X = synth.device_features(hi - lo, F, spec.seed, dev, row_lo=lo)
        labels, mask = synth.device_labels(s_hi - s_lo, C, spec.seed, dev, row_lo=s_lo)
        full = None
        graphs_local, edges = [], 0
        for p in paths:
            ip, ix = synth.device_random_csr(hi - lo, N, spec.mean_degree, spec.seed + 17 * (p + 1), dev, row_lo=lo,
                                             powerlaw=spec.powerlaw)
            graphs_local.append(hb.MetaPathGraph.from_csr(ip, ix, n_cols=N, row_offset=lo))
            edges += graphs_local[-1].nnz
        if dist:
            edges = int(dist.all_reduce_sum(torch.tensor([edges], dtype=torch.float64, device=dev)).item())
        host = None
    if name in synth.SMALL:
        graphs_local = [full[p].row_slice(lo, hi) if dist else full[p] for p in paths]
    ws = X.numel() * 4 + sum(g.nnz * 4 * 3 for g in graphs_local) + (hi - lo) * len(graphs_local) * (72 + 96 + 64 * 3) * 4
    return dict(N=N, F=F, C=C, P=P, X=X, graphs=graphs_local, labels=labels, mask=mask, edges=edges, host=host,
                lo=lo, hi=hi, working_set_bytes=ws, full_graphs=full)


def hb_bias(mask_rows: np.ndarray) -> np.ndarray:
    """Reference-style dense fp32 bias rows (0 on edges, -1e9 elsewhere): what ex_acm3025.py feeds."""
    return np.where(mask_rows, np.float32(0.0), np.float32(-1e9)).astype(np.float32)


def algorithmic_bytes(name, wl):
    """ALGORITHMIC bytes of one launch set (all P meta-paths) of the two gather kernels:
       han_attn_fwd_chunked    : (4 + 4*TS) B/edge + (8 + 4K f1 + 4K lse + 4K c + 3 * 4D [out, V, V']) B/row
       han_attn_bwd_src_chunked: (4 + 4*RS) B/edge + (8 + 4*TS + 4D + 4K) B/source row"""
    K, D = K_HEADS, K_HEADS * HID
    TS, RS = 64, 88       # table rows hold S only (f2 is recomputed from them)
    E = sum(g.nnz for g in wl["graphs"])
    n = wl["hi"] - wl["lo"]
    P = len(wl["graphs"])
    if name in ("han_attn_fwd_chunked", "han_attn_fwd_chunked_split"):
        return (4 + 4 * TS) * E + (8 + 12 * K + 12 * D) * n * P
    if name in ("han_attn_bwd_src_chunked", "han_attn_bwd_src_chunked_split"):
        return (4 + 4 * RS) * E + (8 + 4 * TS + 4 * D + 4 * K) * n * P
    return None


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line only)
    import han_b200 as hb
    from han_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        from han_b200 import synth as _sy
        n_paths = _sy.LARGE[args.workload].P if args.workload in _sy.LARGE else _sy.SMALL[args.workload]().P
        args.partition = pick_partition(args, world, n_paths)
    if world > 1 and args.partition == "tile":
        from han_b200 import tiles as ht
        dist = ht.TileShard.init_process_group(n_paths)
    elif world > 1:
        from han_b200 import dist as hd
        dist = hd.RowShard.init_process_group()
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")

    wl = build_workload(args.workload, dev, dist)
    N, F, C, P = wl["N"], wl["F"], wl["C"], wl["P"]
    gen = torch.Generator(device="cpu").manual_seed(1234)
    hp = hb.HANParams([F] * P, C, (HID,), (K_HEADS, 1), ATT, device=dev, generator=gen)
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    if dist:
        dist.bind(wl["graphs"], N)       # one-off edge exchange (like adj_to_bias, outside the step)
    else:
        for g in wl["graphs"]:
            g.transpose()          # built once per graph (like adj_to_bias, outside the step)
    X1 = wl["X"].unsqueeze(0)
    pmode = projection_mode(args)
    if args.workload in KNOWN_EDGES:
        assert wl["edges"] == KNOWN_EDGES[args.workload], (wl["edges"], KNOWN_EDGES[args.workload])

    def step(Xin, graphs, mask=None, with_l2=True, keep=None):
        """One forward + backward.  mask / with_l2 / keep are used by the parity leg only (loss over a row sample,
        no L2 term, outputs kept)."""
        hp.zero_grad(set_to_none=True)
        logits, fe, av = hb.HeteGAT_multi.inference([Xin] * len(graphs), C, N, True, args.dropout, args.dropout, graphs, [HID], [K_HEADS, 1],
                                                    params=hp, dist=dist, project_mode=pmode)
        msk = wl["mask"] if mask is None else mask
        if dist is None:
            total = hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, C), wl["labels"], msk)
            if with_l2:
                total = total + train.l2_loss()
        else:
            total = dist.masked_loss(logits.reshape(-1, C), wl["labels"], msk, train if with_l2 else None)
        total.backward()
        if dist is not None:
            dist.all_reduce_grads(hp)
        if keep is not None:
            keep.update(final_embed=fe.detach(), att_val=av.detach())
        return total

    flush = None
    if wl["working_set_bytes"] < (1 << 30):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(X1, wl["graphs"])
    barrier()
    # keep the GPU busy until nvidia-smi has produced its first samples, so the timed region is covered
    t_wait = time.perf_counter()
    while rank == 0 and sampler.proc is not None and sampler.mark() < 2 and time.perf_counter() - t_wait < 5.0:
        step(X1, wl["graphs"])
        torch.cuda.synchronize()
    barrier()

    # launch count of one step (eager, counted by the ctypes layer), then optional whole-step graph capture
    counter = _lib.CallRecorder(time_events=False)
    _lib.set_recorder(counter)
    step(X1, wl["graphs"])
    _lib.set_recorder(None)
    launches_per_step = counter.launches
    run_step = lambda: step(X1, wl["graphs"])
    if args.cuda_graph:
        from han_b200.graphs import GraphedStep
        run_step = GraphedStep(run_step, warmup=1)
    barrier()

    mark_lo = sampler.mark()
    nvtx_id = torch.cuda.nvtx.range_start("han_timed")   # start/end ranges span all threads (autograd runs backward in its own)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        if flush is not None:
            flush.zero_()
        s.record()
        loss = run_step()
        e.record()
    barrier()
    torch.cuda.nvtx.range_end(nvtx_id)
    if rank == 0:
        time.sleep(0.25)     # let the sampler flush the samples taken during the region
    clocks = sampler.stop(max(0, mark_lo - 1), None) if rank == 0 else None
    # per-kernel CUDA events: the same K steps again, eagerly, with an event pair around every C-ABI call
    # (graph replays cannot carry timing events; kernel durations do not depend on how they are launched)
    rec = _lib.CallRecorder(time_events=True)
    _lib.set_recorder(rec)
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in ev2:
        if flush is not None:
            flush.zero_()
        s.record()
        step(X1, wl["graphs"])
        e.record()
    barrier()
    _lib.set_recorder(None)
    eager_ms = sum(s.elapsed_time(e) for s, e in ev2) / len(ev2)
    if os.environ.get("HAN_TRACE") and rank == 0:
        # debugging aid: timeline of one eager step (all streams) -> gpurun_out/trace_*.txt
        _lib.TRACE = []
        _lib.trace_mark("step >")
        step(X1, wl["graphs"])
        _lib.trace_mark("step <")
        torch.cuda.synchronize()
        t0 = _lib.TRACE[0][1]
        rows = sorted((t0.elapsed_time(ev), lab) for lab, ev in _lib.TRACE)
        _lib.TRACE = None
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = os.environ.get("HAN_DIST_COMM", "pull") if world > 1 else "single"
        with open(os.path.join(ROOT, "gpurun_out", f"trace_{tag}_g{world}.txt"), "w") as f:
            for t, lab in rows:
                f.write(f"{t:9.3f}  {lab}\n")
    elif os.environ.get("HAN_TRACE"):
        step(X1, wl["graphs"])      # keep the other ranks in lock-step with rank 0's traced step
    barrier()
    step_ms = [s.elapsed_time(e) for s, e in ev]
    ms = sum(step_ms) / len(step_ms)
    loss = loss.detach().clone().reshape(1)
    if dist:
        ms = dist.all_reduce_max(torch.tensor([ms], dtype=torch.float64, device=dev)).item()
        loss = dist.all_reduce_sum(loss)
    value = wl["edges"] / (ms * 1e-3)

    # ---- dominant kernel roofline (rank-local, all ranks do the same work) ---------------------
    summ = rec.summary()
    top = max((k for k in summ if algorithmic_bytes(k, wl)), key=lambda k: summ[k][1])
    calls, tot_ms = summ[top]
    launches_per_set = len(wl["graphs"])       # one launch per local meta-path and step
    sets = calls / launches_per_set
    achieved = algorithmic_bytes(top, wl) * sets / (tot_ms * 1e-3) / 1e9
    peak, peak_src = peaks()
    # DRAM traffic of the dominant kernel from the committed ncu capture -- only when THIS launch has the shape that was
    # profiled (same algorithmic bytes per launch); otherwise null (sharded runs, other workloads)
    traffic = None
    alg_per_launch = int(algorithmic_bytes(top, wl) / launches_per_set)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            ent = json.load(f).get(args.workload, {}).get(top)
        if isinstance(ent, dict) and abs(ent.get("algorithmic_bytes_per_launch", 0) - alg_per_launch) <= 1e-3 * alg_per_launch:
            traffic = ent["dram_bytes_per_launch"]
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": round(tot_ms / calls, 4),
                "algorithmic_bytes_per_launch": alg_per_launch,
                # share of the TIMED step (graph replay: device time only), comparable with the ncu launch list
                "kernel_share_of_step": round((tot_ms / args.steps) / ms, 4),
                "timing": "per-launch CUDA events over K eager steps run right after the timed region",
                "eager_ms_per_step": round(eager_ms, 3),
                "kernels_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in sorted(summ.items())}}

    # ---- end-to-end through the public API with HOST buffers -----------------------------------
    e2e = None if args.no_e2e else run_e2e(args, wl, hp, train, dist, dev, step)
    parity = None if args.no_parity else parity_check(args, wl, hp, dist, dev, step, rank, world)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, world, wl["edges"]),
           "roofline": roofline, "e2e": e2e, "parity": parity, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
           "cuda_graph": bool(args.cuda_graph),
           "loss": float(loss)}
    assert (flush is not None) == (args.workload in _lib_synth().SMALL)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=20.0)
        if args.workload not in _lib_synth().SMALL and not args.no_secondary:
            # like-for-like line: on the ACM3025 shape (BASELINE configs[0]) the COMPLETE dense reference step runs
            # on the host, so both sides are timed on the same config in this same run
            del wl, X1
            torch.cuda.empty_cache()
            out["secondary"] = secondary_acm(dev)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist:
        # Leave without tearing NCCL down: destroying a communicator whose collectives were captured in a
        # CUDA graph can block at interpreter exit; all results are out, all ranks are past the last barrier.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _lib_synth():
    from han_b200 import synth
    return synth


def secondary_acm(dev):
    """ACM3025-shaped graph at full size: GPU step (CUDA-graph replay, L2 flushed) next to the complete dense fp32
    reference step on the host cores (oracle port: adj_to_bias biases, 16 attn_head calls, SimpleAttLayer, dense,
    masked CE + L2, autograd) -- same inputs, same step definition, same run."""
    import han_b200 as hb
    from han_b200 import synth
    from han_b200.graphs import GraphedStep
    cfg = synth.SMALL["acm"]()
    hp = hb.HANParams([cfg.F] * cfg.P, cfg.C, (HID,), (K_HEADS, 1), ATT, device=dev,
                      generator=torch.Generator(device="cpu").manual_seed(1234))
    train = hb.BaseGAttN.training(hp, 0.005, 0.001)
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    for g in graphs:
        g.transpose()
    X = torch.from_numpy(cfg.X).to(dev)[None]
    labels = torch.from_numpy(cfg.labels).to(dev)
    mask = torch.from_numpy(cfg.train_mask.astype(np.float32)).to(dev)

    def step():
        hp.zero_grad(set_to_none=True)
        logits, _, _ = hb.HeteGAT_multi.inference([X] * cfg.P, cfg.C, cfg.N, True, 0.0, 0.0, graphs, [HID], [K_HEADS, 1],
                                                  params=hp, project_mode=2)
        total = hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, cfg.C), labels, mask) + train.l2_loss()
        total.backward()
        return total
    for _ in range(3):
        step()
    run = GraphedStep(step, warmup=1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    torch.cuda.synchronize()
    for s_, e_ in ev:
        flush.zero_()
        s_.record()
        loss = run()
        e_.record()
    torch.cuda.synchronize()
    gpu_ms = sum(s_.elapsed_time(e_) for s_, e_ in ev) / len(ev)
    cpu = cpu_baseline("acm", steps=2, warmup=1)
    cpu_ms = statistics.median(cpu["step_s"]) * 1e3
    return {"workload": WORKLOADS["acm"], "edges": cfg.n_edges(), "gpu_ms_per_step": round(gpu_ms, 4),
            "gpu_value": cfg.n_edges() / (gpu_ms * 1e-3), "cpu_reference_ms_per_step": round(cpu_ms, 2),
            "cpu_reference_value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
            "gpu_over_cpu": round(cpu_ms / gpu_ms, 1), "gpu_loss": float(loss.detach()),
            "note": "complete reference step on both sides (BASELINE configs[0] shape), device-resident inputs"}


def parity_check(args, wl, hp, dist, dev, step, rank, world, rows_per_block=250, blocks=8):
    """Parity evidence inside the driver-run record.  After the timed region, on every N: one extra forward + backward
    whose loss is the cross-entropy over a SAMPLE of destination rows (8 blocks of 250 rows spread over the graph, plus
    every rank's highest-degree row -- the hub rows of the power-law config), and the same quantity from the fp64
    edge-list twin of the reference (oracle/han_oracle.py: inference_edges + autograd) on the sample's receptive field,
    regenerated on rank 0 from the seeded generators.  Compared: final_embed and att_val of the sampled rows, the
    loss, and EVERY all-reduced gradient.  The oracle is the checker here, never the thing measured."""
    import han_b200 as hb
    from han_b200 import synth
    from oracle import han_oracle as O
    N, F, C, P = wl["N"], wl["F"], wl["C"], wl["P"]
    tile = dist if (dist is not None and hasattr(dist, "paths")) else None
    if tile is not None:
        (a_lo, a_hi), (s_lo, s_hi) = tile.rows(N)
    else:
        a_lo, a_hi = wl["lo"], wl["hi"]
        s_lo, s_hi = a_lo, a_hi
    # ---- the sample (identical on every rank) ----
    starts = [int(b * N // blocks) for b in range(blocks)]
    rb = min(rows_per_block, max(1, N // blocks))
    rows = torch.cat([torch.arange(st, min(N, st + rb)) for st in starts])
    g0 = wl["graphs"][0]
    deg = g0.indptr[1:] - g0.indptr[:-1]
    hub = (deg.argmax() + a_lo).reshape(1).to(torch.int64)
    if dist is not None:
        import torch.distributed as td
        hubs = [torch.zeros_like(hub) for _ in range(world)]
        td.all_gather(hubs, hub)
        hub = torch.cat(hubs)
    rows = torch.unique(torch.cat([rows, hub.cpu()]))                 # sorted global row ids
    M = int(rows.numel())
    # ---- product: loss over the sample only, no L2 ----
    local = rows[(rows >= s_lo) & (rows < s_hi)]
    mask = torch.zeros(s_hi - s_lo, dtype=torch.float32, device=dev)
    mask[(local - s_lo).to(dev)] = 1.0
    keep = {}
    if dist is not None:
        dist.bind(wl["graphs"], N)        # the host-fed leg re-bound the exchange structures to its own graph handles
    loss = step(wl["X"].unsqueeze(0), wl["graphs"], mask=mask, with_l2=False, keep=keep).detach().reshape(1).clone()
    D = keep["final_embed"].shape[1]
    buf = torch.zeros(M, D + P, dtype=torch.float32, device=dev)
    pos = torch.searchsorted(rows, local).to(dev)
    buf[pos, :D] = keep["final_embed"][(local - s_lo).to(dev)]
    buf[pos, D:] = keep["att_val"][(local - s_lo).to(dev)]
    if dist is not None:
        dist.all_reduce_sum(buf)
        dist.all_reduce_sum(loss)
    grads = hp.grad_dict()
    torch.cuda.synchronize()
    if rank != 0:
        return None
    # ---- oracle on the receptive field (rank 0) ----
    t0 = time.perf_counter()
    csr_rows = []                     # per meta-path: (counts int64[M], cols int64[nnz]) of the sampled rows
    if args.workload in synth.SMALL:
        cfg = synth.SMALL[args.workload]()
        for m in cfg.masks:
            sub = m[rows.numpy()]
            r, c = np.nonzero(sub)
            csr_rows.append((torch.from_numpy(np.bincount(r, minlength=M)), torch.from_numpy(c.astype(np.int64))))
    else:
        spec = synth.LARGE[args.workload]
        # contiguous runs of sampled rows -> one generator call per run (row r's edges depend on (seed, r) only)
        cut = torch.nonzero(rows[1:] != rows[:-1] + 1).reshape(-1) + 1
        runs = torch.tensor_split(rows, cut.tolist())
        for p in range(P):
            cnts, cols = [], []
            for run in runs:
                ip, ix = synth.device_random_csr(int(run.numel()), N, spec.mean_degree, spec.seed + 17 * (p + 1), dev,
                                                 row_lo=int(run[0]), powerlaw=spec.powerlaw)
                cnts.append((ip[1:] - ip[:-1]).cpu())
                cols.append(ix.cpu().to(torch.int64))
            csr_rows.append((torch.cat(cnts), torch.cat(cols)))
    U = torch.unique(torch.cat([rows] + [c for _, c in csr_rows]))
    rest = U[~torch.isin(U, rows)]
    order = torch.cat([rows, rest])                                    # sampled rows first
    lut = torch.full((N,), -1, dtype=torch.int64)
    lut[order] = torch.arange(order.numel())
    if args.workload in synth.SMALL:
        Xu = torch.from_numpy(cfg.X[order.numpy()]).double()
        labels = torch.from_numpy(cfg.labels[rows.numpy()]).double()
    else:
        Xu = torch.empty(order.numel(), F, dtype=torch.float64)
        ch = 1 << 18
        for c0 in range(0, N, ch):                                     # regenerate X chunk by chunk, keep the needed rows
            sel = (order >= c0) & (order < c0 + ch)
            if bool(sel.any()):
                blk = synth.device_features(min(ch, N - c0), F, spec.seed, dev, row_lo=c0)
                Xu[sel] = blk[(order[sel] - c0).to(dev)].double().cpu()
        labels = synth.device_labels(N, C, spec.seed, dev)[0][rows.to(dev)].double().cpu()
    nU = int(order.numel())
    csr_list = []
    for cnt, cols in csr_rows:
        indptr = np.zeros(nU + 1, dtype=np.int64)
        indptr[1:M + 1] = np.cumsum(cnt.numpy())
        indptr[M + 1:] = indptr[M]
        csr_list.append((indptr, lut[cols].numpy()))
    p64 = O.params_to({k: ([t.detach().double().cpu() for t in v] if isinstance(v, list) else v.detach().double().cpu())
                       for k, v in hp.to_dict().items()}, torch.float64, requires_grad=True)
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    logits_o, fe_o, av_o = O.inference_edges([Xu] * P, csr_list, p64, [K_HEADS, 1], [HID], ATT)
    ce_o = -(labels * torch.log_softmax(logits_o[0, :M], dim=-1)).sum(-1).sum() / M
    ce_o.backward()

    def rel(a, b):
        a, b = a.detach().double().cpu(), b.detach().double()
        den = b.abs().max().item()
        return (a - b).abs().max().item() / (den if den > 0 else 1.0)
    errs = {"final_embed": rel(buf[:, :D], fe_o[:M]), "att_val": rel(buf[:, D:], av_o[:M]), "loss": rel(loss[0], ce_o)}
    for k, v in p64.items():
        if isinstance(v, list):
            for i, t in enumerate(v):
                errs[f"d{k}[{i}]"] = rel(grads[k][i], t.grad)
        else:
            errs[f"d{k}"] = rel(grads[k], v.grad)
    worst = max(errs, key=errs.get)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    return {"max_rel": float(f"{errs[worst]:.3e}"), "worst": worst, "worst6": {k: float(f"{v:.2e}") for k, v in top},
            "rows": M, "receptive_field_nodes": nU,
            "edges_checked": int(sum(int(c.sum()) for c, _ in csr_rows)), "max_row_degree": int(max(int(c.max()) for c, _ in csr_rows)),
            "tolerance": 1e-5, "ok": bool(errs[worst] <= 1e-5), "tensors": len(errs),
            "against": "fp64 edge-list twin of the reference (oracle, pinned to the reference's own source) on the sample's "
                       "receptive field; ||a-b||inf/||b||inf per tensor: final_embed, att_val, loss, every all-reduced gradient",
            "oracle_s": round(time.perf_counter() - t0, 1)}


def run_e2e(args, wl, hp, train, dist, dev, step):
    """Same metric through the reference-facing API with HOST inputs: every step copies that step's
    features and graph structure from pinned host memory, rebuilds the device-side graph handles
    (dense bias -> CSR for the small configs, CSR + transposed view for the large ones), runs
    fwd+bwd and reads the loss back."""
    import han_b200 as hb
    n_e2e = max(1, min(args.steps, 20))
    P = wl["P"]
    copy_s, prep_s = torch.cuda.Stream(), torch.cuda.Stream()
    trace = bool(os.environ.get("HAN_E2E_TRACE"))

    def marker(marks):
        if not trace:
            return lambda label, stream: None

        def mark(label, stream):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            marks.append((label, ev))
        return mark

    # A step is split in two so that consecutive steps pipeline like any host-fed loop: stage(k+1) -- the H2D copies of
    # step k+1's inputs on the copy stream (and, single GPU, the by-source views on a third stream as each graph lands)
    # -- is issued right after compute(k) and runs under its kernels.  Every timed step still performs its own copies
    # and its own device-side graph build; inputs are multi-buffered.  Nothing in stage() waits for the compute
    # stream: the facts the host needs about a graph (empty rows, maximum degree) come from the host arrays or are
    # read back on the staging streams.
    if wl["host"] is not None and wl["host"]["bias"] is not None:
        hX, hB = wl["host"]["X"], wl["host"]["bias"]
        h2d = hX.numel() * 4 + sum(b.numel() * 4 for b in hB)

        def stage(slot=0, after=None):
            main = torch.cuda.current_stream()
            copy_s.wait_stream(main) if after is None else copy_s.wait_event(after)
            with torch.cuda.stream(copy_s):
                X = hX.to(dev, non_blocking=True).unsqueeze(0)
                B = [b.to(dev, non_blocking=True) for b in hB]
                ready = torch.cuda.Event()
                ready.record(copy_s)
            return {"X": X, "B": B, "ready": ready, "marks": []}

        def compute(st):
            main = torch.cuda.current_stream()
            main.wait_event(st["ready"])
            graphs = [hb.MetaPathGraph.from_dense_bias(b) for b in st["B"]]
            out = step(st["X"], graphs)
            for t in [st["X"]] + st["B"]:
                t.record_stream(main)
            st["done"] = torch.cuda.Event()
            st["done"].record(main)
            return out

        def views(st):
            pass
    else:
        tile = dist if (dist is not None and hasattr(dist, "gather_features") and dist.fused_z) else None
        if tile is not None:
            # tile sharding: this rank uploads only its SEMANTIC rows of X (1/W of the matrix); the ranks of a row block
            # hand their slices to one another over NVLink
            (a_lo, _), (s_lo, s_hi) = tile.rows(wl["N"])
            hX = wl["X"][s_lo - a_lo:s_hi - a_lo].cpu().pin_memory()
        else:
            hX = wl["X"].cpu().pin_memory()
        hG = [(g.indptr.cpu().pin_memory(), g.indices.cpu().pin_memory()) for g in wl["graphs"]]
        h2d = hX.numel() * 4 + sum(a.numel() * 8 + b.numel() * 4 for a, b in hG)

        def stage(slot=0, after=None):
            # staging on a copy stream: the graphs first, the features last -- single GPU: each graph's by-source
            # view is built on a third stream as soon as that graph has landed.  The kernels wait per graph / per view
            # (MetaPathGraph.ready).
            main = torch.cuda.current_stream()
            copy_s.wait_stream(main) if after is None else copy_s.wait_event(after)
            marks = []
            mark = marker(marks)
            mark("start", main)
            graphs = []
            for i, (a, b) in enumerate(hG):
                graphs.append(hb.MetaPathGraph.from_csr(a, b, n_cols=wl["N"], device=dev, row_offset=wl["lo"], stream=copy_s))
                mark(f"graph {i} on device", copy_s)
            with torch.cuda.stream(copy_s):
                if tile is not None:
                    X = tile.gather_features(hX, copy_s, slot=slot).unsqueeze(0)
                else:
                    X = hX.to(dev, non_blocking=True).unsqueeze(0)
                x_ready = torch.cuda.Event()
                x_ready.record(copy_s)
                mark("X on device", copy_s)
            return {"X": X, "graphs": graphs, "ready": x_ready, "marks": marks}

        def views(st):
            # single GPU: the by-source views, each on a third stream as soon as its graph has landed (their one
            # device->host read -- the maximum in-degree -- waits for that stream only)
            if not dist:
                mark = marker(st["marks"])
                for i, g in enumerate(st["graphs"]):
                    g.transpose(stream=prep_s)
                    mark(f"by-source view {i} built", prep_s)

        def compute(st):
            main = torch.cuda.current_stream()
            graphs, X = st["graphs"], st["X"]
            mark = marker(st["marks"])
            if dist:
                for g in graphs:
                    g.wait_ready()          # the graphs have landed; the feature matrix may still be on PCIe
                if hasattr(dist, "reset"):
                    dist.reset()
                else:
                    dist._bwd = {}
                dist.bind(graphs, wl["N"])  # the by-source edge exchange of THIS step's graphs (a collective: main stream)
                mark("bind done", main)
            main.wait_event(st["ready"])
            out = step(X, graphs)
            mark("step done", main)
            if tile is None:
                X.record_stream(main)
            st["done"] = torch.cuda.Event()
            st["done"].record(main)
            return out

    def pipeline(n):
        """n steps.  Per step k: queue the copies of step k+1 (non-blocking, so the copy stream never idles), issue
        compute(k), build step k+1's by-source views, read the loss of step k back."""
        walls, host_loss, last_marks, prev_done = [], None, [], None
        staged = stage(0)                   # step 0's inputs: inside the timed region like every other step's
        views(staged)
        for k in range(n):
            t1 = time.perf_counter()
            # the staging streams may start once step k-1 is over (it is: its loss was read), not after step k
            nxt = stage((k + 1) & 1, after=prev_done) if k + 1 < n else None
            loss = compute(staged)
            if nxt is not None:
                views(nxt)
            t2 = time.perf_counter()
            host_loss = float(loss.detach())    # D2H read of the step's result (synchronises the step)
            walls.append((round((t2 - t1) * 1e3, 1), round((time.perf_counter() - t2) * 1e3, 1)))
            last_marks, prev_done = staged["marks"], staged["done"]
            staged = nxt
        return walls, host_loss, last_marks

    pipeline(3)                             # warm-up with the same buffering pattern (allocator, pinned staging)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    # what the host link gives this process for a plain pinned copy (explains the e2e floor: h2d bytes / this)
    probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty_like(probe_h, device=dev)
    probe_d.copy_(probe_h, non_blocking=True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        probe_d.copy_(probe_h, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * probe_h.numel() / (s.elapsed_time(e) * 1e-3) / 1e9
    del probe_h, probe_d
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    walls, host_loss, last_marks = pipeline(n_e2e)
    e.record()
    torch.cuda.synchronize()
    ms = max(s.elapsed_time(e), (time.perf_counter() - t0) * 1e3) / n_e2e
    if trace:
        sys.stderr.write(f"e2e per step (host issue ms, wait-for-loss ms): {walls}; events {s.elapsed_time(e):.1f} ms\n")
        if last_marks:
            t_0 = last_marks[0][1]
            sys.stderr.write("e2e timeline of the last step (ms): " +
                             ", ".join(f"{lab} {t_0.elapsed_time(ev):.1f}" for lab, ev in last_marks[1:]) + "\n")
    if dist:
        ms = dist.all_reduce_max(torch.tensor([ms], dtype=torch.float64, device=dev)).item()
    return {"value": wl["edges"] / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 4, "ms_per_step": ms, "steps": n_e2e, "loss": host_loss,
            "pipelining": "multi-buffered inputs: the H2D copies and graph build of step k+1 run under the kernels of step k",
            "h2d_link_gbs": round(h2d_gbs, 1), "h2d_floor_ms": round(h2d / (h2d_gbs * 1e9) * 1e3, 1)}


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_baseline(workload, budget_s=20.0, steps=None, warmup=1):
    from han_b200 import synth
    from oracle import cpu_reference as cr
    from oracle import han_oracle as O
    cores = cr.host_threads()
    if workload in synth.SMALL:
        cfg = synth.SMALL[workload]()
        params = O.init_params(np.random.default_rng(1), [cfg.F] * cfg.P, cfg.C, dtype=torch.float32)
        times = cr.dense_full_step_seconds(cfg, params, steps or 2, warmup)
        t = statistics.median(times)
        return {"value": cfg.n_edges() / t, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{len(times)} complete dense fp32 reference steps (fwd+bwd, all {cfg.P * K_HEADS} heads) of the "
                          f"{workload} config on torch-CPU; median {t:.3f} s/step", "step_s": times}
    spec = synth.LARGE[workload]
    R = 16
    prob = cr.make_rowblock_problem(spec.N, spec.P, K_HEADS, HID, R, min(spec.mean_degree, 64), 7)
    times, edges = [], cr.rowblock_edges(prob)
    t_all = time.perf_counter()
    for it in range((steps or 2) + warmup):
        dt = cr.dense_rowblock_step_seconds(prob, K_HEADS, HID)
        if it >= warmup:
            times.append(dt)
        if steps is None and time.perf_counter() - t_all > budget_s:
            break
    times = times or [dt]
    t = statistics.median(times)
    return {"value": edges / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"dense chain (utils/layers.py:26-35,46) fwd+bwd for {R} of {spec.N} destination rows x all "
                      f"{spec.N} source columns, {spec.P} meta-paths x {K_HEADS} heads, torch-CPU fp32; projection/"
                      f"semantic excluded (favours CPU); N x N does not fit host memory; median {t:.2f} s per sample",
            "step_s": times}


def run_reference(args):
    """--impl reference: the reference's dense CPU algorithm (oracle port; TF1 cannot be installed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(args.workload, steps=args.steps, warmup=args.warmup)
    step_s = statistics.median(base["step_s"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    edges = None
    from han_b200 import synth
    if args.workload in synth.LARGE and args.workload not in KNOWN_EDGES:
        # count the edges of the configured graph on the host (seeded generator, chunk by chunk)
        spec = synth.LARGE[args.workload]
        edges = 0
        for p_ in range(spec.P):
            for c0 in range(0, spec.N, 1 << 16):
                ip, _ = synth.device_random_csr(min(1 << 16, spec.N - c0), spec.N, spec.mean_degree,
                                                spec.seed + 17 * (p_ + 1), "cpu", row_lo=c0, powerlaw=spec.powerlaw)
                edges += int(ip[-1])
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": len(base["step_s"]), "warmup": args.warmup, "ms_per_step": step_s * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, world, edges),
           "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="syn2m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row parity check against the fp64 oracle")
    ap.add_argument("--no-secondary", action="store_true", help="skip the ACM-shaped like-for-like line")
    ap.add_argument("--partition", choices=["auto", "row", "tile"], default=os.environ.get("HAN_DIST_PARTITION", "auto"),
                    help="multi-GPU partitioning: destination-row shards, (meta-path x row-block) tiles, or auto "
                         "(tiles when there are more ranks than meta-paths)")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="attn_drop = ffd_drop (the reference trains with 0.6); 0 = parity / headline setting")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="launch every kernel from Python instead of replaying the captured step")
    ap.add_argument("--projection", default="auto", choices=["auto", "fp32", "tf32x3", "tf32x2", "tf32"],
                    help="K-A arithmetic: auto = tcgen05 3xTF32 (real-valued X) / 2xTF32 (0/1 features), both FP32-grade")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                      # W >= 3 for both arms
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
