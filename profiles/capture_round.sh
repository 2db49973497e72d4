#!/bin/bash
# One GPU-box pass that produces a round's evidence under gpurun_out/ (copied to profiles/ afterwards):
#   bench lines (own arm, reference arm), the ncu launch list of the bench command and one `--set full` capture per hot kernel.
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
# usage (on the box): bash profiles/capture_round.sh r2_v1
TAG=${1:-round}
K2ONLY="--k3-reads 0 --k5-steps 0 --no-labels --cpu-sample 0 --preheat 0"
mkdir -p gpurun_out
python bench.py --steps 40 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 2 --warmup 1 $K2ONLY > gpurun_out/${TAG}_bench_short.json 2>> gpurun_out/${TAG}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 $K2ONLY > gpurun_out/${TAG}_ncu_launches.log 2>&1
for k in block_mlp_kernel longconv_tc2_kernel block_in_kernel score_pool_kernel embed_in_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 6 -c 1 -f -o gpurun_out/${TAG}_$k \
      python bench.py --steps 2 --warmup 1 $K2ONLY > gpurun_out/${TAG}_ncu_$k.log 2>&1
  # summarise on the box and drop the 12-14 MB report: gpurun brings back at most 64 MiB
  python profiles/ncu_summary.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_$k.txt 2>&1 && rm -f gpurun_out/${TAG}_$k.ncu-rep
done
tail -c 400 gpurun_out/${TAG}_bench.json
# the chunked form of the tensor-core conv (reads longer than 8 200 tokens): a forward of 16 x 32 769 tokens only
cat > gpurun_out/_long_fwd.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict
eng = Engine(make_state_dict(0), device=0, max_batch=16, max_tokens=32769)
ids = torch.randint(7, 11, (16, 32769), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.forward(ids)
torch.cuda.synchronize()
PY
python gpurun_out/_long_fwd.py && ncu --set full --clock-control none --import-source on -k regex:longconv_tc2_kernel --launch-skip 5 -c 1 -f \
    -o gpurun_out/${TAG}_longconv_tc_chunked python gpurun_out/_long_fwd.py > gpurun_out/${TAG}_ncu_longconv_tc_chunked.log 2>&1
python profiles/ncu_summary.py gpurun_out/${TAG}_longconv_tc_chunked.ncu-rep > gpurun_out/${TAG}_longconv_tc2_chunked.txt 2>&1 && rm -f gpurun_out/${TAG}_longconv_tc_chunked.ncu-rep
