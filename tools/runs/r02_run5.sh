#!/bin/bash
# round-2 GPU run 5: re-validate (graph bit-identity, per-layer yardstick), ncu of the bandwidth kernels
cd "$(dirname "$0")/../.."
O=gpurun_out/run5; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_sdxl_parity_gpu.py > $O/pytest.log 2>&1; tail -6 $O/pytest.log
echo "== sdxl parity"; timeout 1200 python -m pytest tests/test_sdxl_parity_gpu.py -q -s > $O/sdxl_parity.log 2>&1; tail -6 $O/sdxl_parity.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== ncu bw kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_|gn_|geglu|noise_fwd|wmse" -o $O/bw -f python tools/profile_one.py bw > $O/ncu_bw.log 2>&1; tail -2 $O/ncu_bw.log
ls -la $O/*.ncu-rep
echo DONE
