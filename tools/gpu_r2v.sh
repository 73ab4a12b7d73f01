# pentadiagonal solve: parity/accuracy tests + timing against the tridiagonal solve at 4096 x 2048
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pentadiagonal or cosine or fdm or rayleigh" 2>&1 | tail -3
python - <<'PY'
import sys, ctypes, torch
sys.path.insert(0, "melvin.py_b200")
from melvin import _backend
ctx = _backend.Context(4096, 2048, 2.44, 1.0, True, 4)
rhs = torch.randn(ctx.spec_shape, dtype=torch.complex128, device="cuda"); out = torch.empty_like(rhs)
vp = ctypes.c_void_p
for name in ("mlv_solve_fdm", "mlv_solve_fdm_o4"):
    for _ in range(3): ctx.call(name, vp(rhs.data_ptr()), vp(out.data_ptr()))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ctx.call(name, vp(rhs.data_ptr()), vp(out.data_ptr()))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"{name}: {ms:.4f} ms, {2 * rhs.numel() * 16 / ms / 1e6:.0f} GB/s of 2 S_f")
PY
