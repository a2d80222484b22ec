"""bench.py's CPU legs need no GPU: the reference arm (`--impl reference`: the oracle port timed on the host
cores) must keep printing ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "acm",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "han_fwd_bwd_metapath_edges_per_s" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "acm" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_algorithmic_bytes_match_the_design_table():
    """260 B/edge + 872 B/row forward, 356 B/edge + 552 B/source backward (K=8, D=64)."""
    sys.path.insert(0, ROOT)
    import bench

    class G:
        nnz = 1000
    wl = {"graphs": [G(), G()], "lo": 0, "hi": 50, "P": 2}
    assert bench.algorithmic_bytes("han_attn_fwd_chunked", wl) == 260 * 2000 + 872 * 50 * 2
    assert bench.algorithmic_bytes("han_attn_bwd_src_chunked_split", wl) == 356 * 2000 + 552 * 50 * 2
    assert bench.algorithmic_bytes("han_semantic_fwd", wl) is None
