DICP_SMALL_MID=0 python scripts/experiments/mid_case.py 2>&1 | tail -2
DICP_SMALL_MID=1 python scripts/experiments/mid_case.py 2>&1 | tail -2
python - <<'PY'
import numpy as np
for k in range(2):
    a=np.load(f"gpurun_out/mid_case_0_{k}.npy"); b=np.load(f"gpurun_out/mid_case_1_{k}.npy")
    print("old batched vs mid, frame", k, np.abs(a-b).max()/np.abs(a).max())
PY
