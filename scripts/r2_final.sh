#!/usr/bin/env bash
# Round 2, closing run on one B200: smoke, GPU suite, the default bench line, the reference arm.
set -u
O=gpurun_out; mkdir -p $O
echo "== [1] smoke + pytest -m gpu"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/r2_pytest_gpu.log
echo "== [2] default bench"; timeout 900 python bench.py --e2e-stages > $O/r2_bench_10m_n1.json 2> $O/r2_bench_10m_n1.err; tail -2 $O/r2_bench_10m_n1.err
echo "== [3] bench 1M"; timeout 600 python bench.py --points 1000000 --e2e-stages > $O/r2_bench_1m_n1.json 2> $O/r2_bench_1m_n1.err; tail -2 $O/r2_bench_1m_n1.err
echo "== [4] reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; tail -2 $O/r2_bench_reference.err
python - <<'PY'
import json
for f in ("r2_bench_10m_n1", "r2_bench_1m_n1", "r2_bench_reference"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read())
        e = d.get("e2e") or {}
        a = d.get("affinity_kernel") or {}
        print(f, "value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 4), (d.get("roofline") or {}).get("frac"), "e2e", e.get("seconds") or e.get("value"),
              e.get("stages_seconds_rank0_instrumented_refit"), "shuffled e2e", (e.get("shuffled_rows") or {}).get("seconds"), "knn", a.get("ms"), "knn shuffled", (a.get("shuffled_rows") or {}).get("ms"),
              "full TF", a.get("full_sweep_tflops_2nnd"), (d.get("parity") or {}).get("sigma_rows_vs_oracle"), (d.get("parity") or {}).get("knn_sampled_rows_fp64", {}).get("mismatches_on_decided_entries") if d.get("parity") else None,
              (d.get("cpu_baseline") or {}).get("value_measured"), d.get("value_measured"), d.get("clocks"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
