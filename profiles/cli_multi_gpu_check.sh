set -e
cd /root/repo
python - <<'PY'
import numpy as np
from chimeralm_b200.bam import BamWriter, make_record, minimal_header
rng = np.random.default_rng(1)
w = BamWriter("/tmp/cli.bam", minimal_header())
acgt = np.frombuffer(b"ACGT", np.uint8)
for i in range(1000):
    L = int(rng.integers(500, 9000))
    w.write(make_record(f"read_{i:05d}", acgt[rng.integers(0, 4, L)].tobytes(), sa_tag=(i % 10 != 0)))
w.close()
PY
rm -rf /tmp/p1 /tmp/p2 /tmp/p2b
time python -m chimeralm_b200 predict /tmp/cli.bam -o /tmp/p1 -b 32 --gpus 1 2>&1 | grep -v "^│\|^╭\|^╰" | tail -4
time python -m chimeralm_b200 predict /tmp/cli.bam -o /tmp/p2 -b 32 --gpus 2 2>&1 | grep -v "^│\|^╭\|^╰" | tail -4
time python -m chimeralm_b200 predict /tmp/cli.bam -o /tmp/p2b -b 32 --gpus 2 --bucket 2>&1 | grep -v "^│\|^╭\|^╰" | tail -4
python - <<'PY'
from chimeralm_b200.callbacks import load_predictions_from_folder
a = load_predictions_from_folder("/tmp/p1"); b = load_predictions_from_folder("/tmp/p2"); c = load_predictions_from_folder("/tmp/p2b")
print(len(a), len(b), len(c), "1gpu==2gpu:", a == b, "bucketed agree:", sum(a[k] == c[k] for k in a), "/", len(a))
PY
