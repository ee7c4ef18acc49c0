python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import torch, numpy as np
from upmix_b200 import _native, filter_design as fd
n=3600*48000
x=torch.randn(n,device='cuda'); taps=torch.from_numpy(fd.design_lr4_hp_fir(48000,180.0,1025)).cuda()
for _ in range(2): y=_native.fir_filter(x,taps)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(3): y=_native.fir_filter(x,taps)
b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b)/3
print(f"FIR 1025 taps, 1-hour mono track: {ms:.2f} ms -> {2*1025*n/ms/1e9:.1f} TFLOP/s")
PY
