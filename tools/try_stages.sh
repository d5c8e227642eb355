# rebuild with 3 stages per pair and sweep the stage size (same total shared memory as 2 x 4224 at 2816)
python - <<'PY'
from rappas_b200 import build as b
b.build(force=True, extra=["-DRP_STAGES=3"])
PY
SWEEP="${SWEEP3:-3072 3584 4096}" bash tools/sweep_stage.sh
python -m rappas_b200.build --force > /dev/null
