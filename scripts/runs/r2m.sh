#!/bin/bash
# round 2, GPU call M (1 GPU): K2 input queue timing, then the driver's sequence again (smoke, full suite, bench)
set -u
mkdir -p gpurun_out
{
echo "== K2 queue f32"; python scripts/dev_bench_nuts.py f32 2>&1 | tail -2
echo "== K2 queue f64"; python scripts/dev_bench_nuts.py f64 2>&1 | tail -2
echo "== K2 queue f32 normal"; python scripts/dev_bench_nuts.py f32 16384 normal 2>&1 | tail -2
echo "== K2 no queue f32"; CTX_B200_NVRTC_FLAGS="-DCTX_LL_QUEUE=0" CTX_B200_NO_DISK_CACHE=1 python scripts/dev_bench_nuts.py f32 2>&1 | tail -2
} > gpurun_out/r2m_k2.log 2>&1
cat gpurun_out/r2m_k2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2m_smoke.log | cut -c1-200
python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
python bench.py > gpurun_out/r2m_bench.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.log 2>&1; echo "ref rc=$?"
