#!/bin/bash
# round 2, call v (1 GPU): packet preparation without memset / memcpy on the stream; NUMA placement of the pinned messages
mkdir -p gpurun_out
timeout 120 python scratch/numa_h2d_probe.py 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_fe.py tests/test_gpu_fe_pipeline.py tests/test_gpu_stream_device.py tests/test_gpu_firstparty.py -x -q 2>&1 | tail -4
timeout 300 python scratch/e2e_timeline.py 2>&1 | tail -9
for v in "CMAXB_E2E_TRACE=1" "CMAXB_E2E_TRACE=1 CMAXB_E2E_SKIP=eval"; do
  echo "== $v"
  env $v timeout 600 python bench.py --skip-configs --steps 100 > gpurun_out/r2v.json 2> gpurun_out/r2v.err
  grep "e2e host" gpurun_out/r2v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2v.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.3e e2e %.3e (%.1f us/step)" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3))
PY
done
