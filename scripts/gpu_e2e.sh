#!/bin/bash
# Host-side 2-bit packer in front of the H2D copy: parity, packer speed on the box's cores, e2e with / without it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/e2e.log) 2>&1
nproc; grep -m1 "model name" /proc/cpuinfo
echo "=== pytest -m gpu ==="; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== packer speed ==="
python - <<'PY'
import time, numpy as np, sys, ctypes as C
sys.path.insert(0, '.')
import acwm_pkg
acwm = acwm_pkg.load(); L = acwm.lib()
t = np.random.default_rng(1).integers(0, 4, 128 << 20, dtype=np.uint8)
out = np.zeros(32 << 20, np.uint8); bad = C.c_int()
for i in range(6):
    t0 = time.perf_counter(); L.acwm_pack_text_2bit(C.c_void_p(t.ctypes.data), t.size, C.c_void_p(out.ctypes.data), C.byref(bad)); dt = time.perf_counter() - t0
    print(f"pack 128 MiB: {dt*1e3:.2f} ms = {t.size/dt/1e9:.1f} GB/s")
PY
for hp in 1 0; do for wl in c2 c1; do
echo "=== bench $wl ACWM_HOST_PACK=$hp ==="
ACWM_HOST_PACK=$hp timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload $wl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',d['e2e'])"
done; done
