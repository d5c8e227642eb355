# Diagnoses the collapse of the partitioned DB beyond two GPUs.  Needs >= 2 GPUs.
# Hypothesis A (translation reach of the CUDA-IPC mappings): a two-GPU run whose REMOTE posting blocks are as
# large as the four-GPU run's (1.7 GB instead of 1.15 GB) collapses too.  Hypothesis B (several destinations):
# it does not, and only the number of peers matters.
run() { # gpus extra-args tag
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 296$1$RANDOM \
    bench.py --gpus $1 --config 5 --reads 300000 --partitioned --replicate-table --steps 3 --warmup 2 --no-cpu --no-e2e $2 \
    > gpurun_out/diag_part_$3.json 2> gpurun_out/diag_part_$3.err
  grep -h -o '"value": [0-9.e+]*, "unit": "reads/s", "n_gpus": [0-9]*[^}]*"ms_per_step": [0-9.]*' gpurun_out/diag_part_$3.json | sed "s/^/$3: /"
}
N=${N:-2}
run $N "" n${N}_x1
run $N "--postings-scale 1.5" n${N}_x1.5
run $N "--postings-scale 2" n${N}_x2
# the same GPUs in ONE process (peer access enabled by the library, no CUDA-IPC mappings): cfg2-like DB, layout 2
python - <<'PY'
import time
import numpy as np
import rappas_b200 as R
from rappas_b200 import synth
nd = R.device_count()
db = synth.make_db(0, 10, 1999, n_keys=786432, mean_postings=32, seed=44)
rb = synth.make_reads(db, 400000, 150, seed=1044)
for layout, devs in ((0, (0,)), (2, tuple(range(nd)))):
    g = R.Database.from_synth(db, devices=devs, partitioned=layout)
    g.place(rb)
    t0 = time.perf_counter(); g.place(rb); dt = time.perf_counter() - t0
    print("in-process layout %d on %d GPU(s): %.1f M reads/s end to end" % (layout, len(devs), rb.n_reads / dt / 1e6))
    g.close()
PY
python - <<'PY'
import torch
n = torch.cuda.device_count()
print("peer access", [[int(torch.cuda.can_device_access_peer(a, b)) if a != b else 1 for b in range(n)] for a in range(n)])
PY
nvidia-smi topo -m | head -12
