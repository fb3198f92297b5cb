"""Developer check run on the GPU box: reference CUDA build vs CPU oracle vs libpfdtd_b200.

Writes gpurun_out/golden/*.npz (responses + node bytes produced by the reference itself) that
tools/make_golden.py turns into the committed fixtures under tests/golden/.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import casefile, oracle  # noqa: E402
from parallelfdtd_b200 import capi, synth  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(OUT, "golden"), exist_ok=True)
LAM = 1.0 / np.sqrt(3.0)


from tests.fdtd_cases import (make_case, parity_cases, source_table, run_oracle, rel_l2, node_checksum)  # noqa: E402
from tests import fdtd_cases as fc  # noqa: E402


def run_ours(case, n_parts=None, kernel=capi.KERNEL_AUTO, matidx=1, opts=()):
    return fc.run_ours(capi, case, n_parts=n_parts, kernel=kernel, matidx=matidx, opts=opts)


def main():
    results = []
    have_ref = casefile.ref_available()
    print("devices:", capi.device_count(), "ref:", have_ref, flush=True)
    cases = parity_cases()
    for case in cases:
        name = case["name"]
        t0 = time.time()
        r_or, (pos, mat, air, bnd), osecs = run_oracle(case)
        rec = dict(case=name, oracle_s=round(time.time() - t0, 2))
        r_ref = None
        if have_ref:
          try:
            ref = casefile.run_reference(case, os.path.join("/tmp/pfdtd_ref", name), dump_nodes=True)
            r_ref = ref["responses"]
            rec["ref_vs_oracle_bitexact"] = bool(np.array_equal(r_ref, r_or))
            rec["ref_vs_oracle_rel_l2"] = rel_l2(r_or, r_ref)
            first = [p[0] for p in ref["partitions"]]
            ok_nodes = True
            for (f, sz), (rp, rm) in zip(ref["partitions"], ref["nodes"]):
                ok_nodes &= bool(np.array_equal(rp, pos[f:f + sz]) and np.array_equal(rm, mat[f:f + sz]))
            rec["ref_nodes_vs_oracle_bitexact"] = ok_nodes
            rec["ref_counts"] = [ref["n_air"], ref["n_boundary"]]
            rec["oracle_counts"] = [air, bnd]
            np.savez_compressed(os.path.join(OUT, "golden", name + ".npz"), responses=r_ref, dims=np.asarray(ref["dims"]),
                                n_air=ref["n_air"], n_boundary=ref["n_boundary"],
                                partitions=np.asarray(ref["partitions"]),
                                pos_crc=np.asarray([node_checksum(p[0]) for p in ref["nodes"]], dtype=np.uint64),
                                mat_crc=np.asarray([node_checksum(p[1]) for p in ref["nodes"]], dtype=np.uint64),
                                **{f"pos_{k}": p[0] for k, p in enumerate(ref["nodes"])},
                                **{f"mat_{k}": p[1] for k, p in enumerate(ref["nodes"])})
          except Exception as e:  # noqa: BLE001
            rec["ref_error"] = repr(e)
        for kname, kern in (("plain", capi.KERNEL_PLAIN), ("tma", capi.KERNEL_TMA)):
            try:
                r, nodes, info = run_ours(case, kernel=kern)
                rec[kname + "_kernel"] = info["kernel"]
                rec[kname + "_vs_oracle_bitexact"] = bool(np.array_equal(r, r_or))
                rec[kname + "_vs_oracle_rel_l2"] = rel_l2(r, r_or)
                f_, s_ = oracle.partition_indexing(pos.shape[0], case["n_parts"])
                okn = all(np.array_equal(nodes[k][0], pos[f_[k]:f_[k] + s_[k]]) and np.array_equal(nodes[k][1], mat[f_[k]:f_[k] + s_[k]])
                          for k in range(case["n_parts"]))
                rec[kname + "_nodes_bitexact"] = bool(okn)
                rec[kname + "_counts"] = list(info["counts"])
                if r_ref is not None:
                    rec[kname + "_vs_ref_bitexact"] = bool(np.array_equal(r, r_ref))
                # partition invariance
                inv = True
                for n in (1, 2, 5):
                    rn, _, _ = run_ours(case, n_parts=n, kernel=kern)
                    inv &= bool(np.array_equal(rn, r))
                rec[kname + "_partition_invariant_1_2_5"] = inv
            except Exception as e:  # noqa: BLE001
                rec[kname + "_error"] = repr(e)
        print(json.dumps(rec), flush=True)
        results.append(rec)
    json.dump(results, open(os.path.join(OUT, "gpu_check.json"), "w"), indent=1)

    # quick timing at 256^3 / 512^3
    for dims, steps in (((256, 256, 256), 50), ((512, 512, 512), 30)):
        bid, mat = synth.shoebox(dims, 6)
        tab = synth.material_table(list(np.linspace(0.99, 0.5, 6)))
        for dt, dname in ((capi.F32, "f32"), (capi.F64, "f64")):
            for ut in (0, 2):
                for kern, kname in ((capi.KERNEL_PLAIN, "plain"), (capi.KERNEL_TMA, "tma")):
                    tiles = [0] if kern == capi.KERNEL_PLAIN else [1, 2, 3, 4, 5, 6]
                    for tile in tiles:
                        try:
                            s = capi.Solver()
                            s.set_option(capi.OPT_KERNEL, kern)
                            s.set_option(capi.OPT_TMA_TILE, tile)
                            s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0)
                            s.set_option(capi.OPT_TIME_KERNELS, 1)
                            prm = oracle.params(LAM, 0, dt == capi.F64)
                            s.setup_mesh(bid, mat, (32, 4, 1), ut, dt, prm, tab)
                            s.make_partition(1, [0])
                            c = dims[0] // 2
                            src = np.zeros((1, steps * 2), dtype=s.np_dtype)
                            src[0, 1] = 1
                            s.set_sources([[c, c, c]], [0], src)
                            s.set_receivers([[c + 5, c, c]])
                            s.enqueue_steps(0, 5)
                            s.sync()
                            s.enqueue_steps(5, steps)
                            s.sync()
                            tot, kms, nk = s.last_timing()
                            nvox = dims[0] * dims[1] * dims[2]
                            bpv = 13 if dt == capi.F32 else 25
                            print(json.dumps(dict(bench=dims, dtype=dname, update_type=ut, kernel=s.kernel_name(),
                                                  ms_per_step=tot / steps, kernel_ms=kms / max(nk, 1),
                                                  mvox_s=nvox * steps / (tot * 1e-3) / 1e6,
                                                  kernel_gbs=nvox * bpv / (kms / max(nk, 1) * 1e-3) / 1e9)), flush=True)
                            s.close()
                        except Exception as e:  # noqa: BLE001
                            print(json.dumps(dict(bench=dims, dtype=dname, update_type=ut, kernel=kname, tile=tile, error=repr(e))), flush=True)
        if have_ref:
            for ut in (0, 2):
                for dbl in (False, True):
                    case = dict(bid=bid, mat=mat, block=(32, 4, 1), update_type=ut, double=dbl, steps=steps, octave=0, n_parts=1,
                                devices=[0], materials=tab, sources=[(dims[0] // 2,) * 3 + (0, 0, 0)],
                                receivers=[(dims[0] // 2 + 5, dims[0] // 2, dims[0] // 2)], input_data=[])
                    ref = casefile.run_reference(case, "/tmp/pfdtd_ref/bench")
                    print(json.dumps(dict(bench=dims, ref=True, update_type=ut, double=dbl, wall=ref["wall_seconds"],
                                          mvox_s=dims[0] * dims[1] * dims[2] * steps / ref["wall_seconds"] / 1e6,
                                          stdout=ref["stdout"].strip())), flush=True)


if __name__ == "__main__":
    main()
