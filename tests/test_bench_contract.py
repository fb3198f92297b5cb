"""CPU: bench.py's JSON contract where it can be exercised without a GPU -- the reference arm on the CPU oracle port
(--ref-kind port) at config-1 size -- and the product arm's refusal to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_port_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--ref-kind", "port", "--workload", "c1", "--steps", "30", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.strip().splitlines()) == 1            # stdout carries the JSON line and nothing else
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "Mvox-updates/s" and d["unit"] == "Mvox/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--ref-kind", "port", "--workload", "c1", "--gpus", "2", "--steps", "5"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_device(capi):
    if capi.device_count() > 0:
        return
    r = subprocess.run([sys.executable, BENCH, "--workload", "c1", "--steps", "5"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_bench_helpers():
    """host-side pieces of bench.py: receiver placement, the roofline denominator and the committed ncu traffic table"""
    sys.path.insert(0, ROOT)
    import bench
    for dims in [(512, 512, 512), (512, 512, 4096), (64, 64, 64)]:
        rec = bench.receiver_positions(dims)
        cx, cy, cz = dims[0] // 2, dims[1] // 2, dims[2] // 2
        assert len(rec) == 4 and all(0 < x < dims[0] - 1 and 0 < y < dims[1] - 1 and 0 < z < dims[2] - 1 for x, y, z in rec)
        near = rec[0]
        assert abs(near[0] - cx) + abs(near[1] - cy) + abs(near[2] - cz) < 30          # reached within the warm-up at any height
        assert len({r[2] for r in rec}) == 4                                             # spread over the height
    peak, src = bench.measured_peak()
    assert 3000 < peak < 9000 and ("measured" in src or "fallback" in src)
    for dtype, dif in [("f32", 2), ("f64", 2), ("f32", 0), ("f64", 0)]:
        t = bench.ncu_traffic("c2", dtype, 0, dif)
        algo = 133693440 * bench.ALGO_BYTES[dtype]
        assert t is not None and 0.95 * algo < t < 1.06 * algo                          # the kernels move the algorithmic bytes
    assert bench.ncu_traffic("c2", "f32", 2, 3) is None
    assert bench.ALGO_BYTES == {"f32": 13, "f64": 25}


def test_committed_bench_lines_are_well_formed():
    """profiles/r01_bench_*.json: the lines the measurements in profiles/README.md come from carry every contract key,
    a consistent roofline fraction and no slowdown reason."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*.json")))
    assert len(files) >= 10
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches"):
            assert k in d, (f, k)
        if os.path.basename(f) == "r01_bench_f32_dif2_512.json":           # the headline line: default invocation, nothing switched off
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
            assert d["roofline"]["traffic"] and len(d["variants"]) >= 6
        assert d["metric"] == "Mvox-updates/s" and d["unit"] == "Mvox/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["value"] > 0 and d["steps"] >= 1
        if d.get("impl") == "reference":
            assert d["cpu_baseline"]["kind"] in ("reference", "port")
            continue
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["e2e"] is None or {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
        assert not bad, (f, bad)


def test_round2_bench_lines_carry_the_new_blocks():
    """profiles/r02_*: median-of-blocks timing, the like-for-like block with its own roofline and e2e, the config-4 block,
    and -- at N > 1 -- the single-GPU invariance check and the halo figures."""
    def load(rel):
        return json.loads(open(os.path.join(ROOT, "profiles", rel)).read().strip().splitlines()[-1])
    n1, n2, n8 = load("r02_bench_n1.json"), load("r02_bench_n2_peer_stores.json"), load(os.path.join("r02_n8", "bench_n8_auto.json"))
    for d in (n1, n2, n8):
        assert len(d["step_ms_blocks"]) == 5 and sorted(d["step_ms_blocks"])[2] == d["ms_per_step"]
        l = d["like_for_like"]
        assert l["value"] > d["value"] and l["roofline"]["frac"] > 0.9 and l["e2e"]["value"] > 0
        assert "frequency-independent" in l["config"]["boundaries"] and "frequency-dependent" in d["config"]["boundaries"]
        assert d["e2e"]["setup_seconds"] > 0 and d["e2e"]["run_seconds"] > 0 and len(d["e2e"]["all_seconds"]) >= 3
        assert d["roofline"]["kernel_ms_per_launch"] <= d["ms_per_step"] * 1.01        # a launch never takes longer than the step
        assert not d["clocks"]["reasons"] or d["clocks"]["reasons"] == ["sw_power_cap"]
        assert d["c4"]["overlap"]["value"] > 0 and "1024x1024x960" in d["c4"]["workload"]
    assert n1["roofline"]["frac"] > 0.9 and n1["gpu_launches"] >= 21
    for d in (n2, n8):
        assert d["slab_invariance"] is True and d["slab_invariance_detail"]["headline"]["max_abs_diff"] == 0.0
        assert "peer-mapped stores" in d["config"]["halo"] and d["halo_ms_per_exchange_alone"] > 0
    # weak scaling at the stated size: config 4, 1.007e9 voxels per GPU, 8.05e9 at N = 8
    eff = n8["c4"]["overlap"]["value"] / (8 * n1["c4"]["overlap"]["value"])
    assert "8.05e9" in n8["c4"]["workload"] and eff > 0.85, eff
    # configs 3 and 5 at their stated size on 8 GPUs
    c3 = json.loads(open(os.path.join(ROOT, "profiles", "r02_n8", "configs_c3_n8_final.jsonl")).read().strip().splitlines()[0])
    assert c3["voxels"] == 1536 * 1024 * 960 and c3["n_gpus"] == 8 and c3["receivers_reached"] == [16] and c3["same_responses_with_nccl_transport"]
    c5 = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r02_n8", "configs_c3_c5_n8.jsonl")).read().strip().splitlines()][-1]
    assert c5["config"] == "c5" and c5["voxels"] == 2048 * 1024 * 1920 and c5["dtype"] == "f64" and c5["distinct_responses"] == 10
