"""Times the tensor-path coarse kernel (CUDA events inside the library) for a synthetic store.
usage: python tools/time_tensor.py rows dim storage k batch"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from archi_b200.store import NativeStore

rows, dim, storage, k, batch = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
s = NativeStore(dim, "cosine", storage, capacity_rows=rows)
for st in range(0, rows, 262144):
    m = min(262144, rows - st)
    x = torch.randn((m, dim), generator=g, device=dev)
    s.append(x / x.norm(dim=1, keepdim=True))
q = torch.randn((batch, dim), generator=g, device=dev)
q = q / q.norm(dim=1, keepdim=True)
for _ in range(3):
    s.search(q, k, path=2)
s.set_timing(True)
ms = []
for _ in range(5):
    s.search(q, k, path=2)
    ms.append(s.last_stats().last_kernel_ms)
s.set_timing(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    s.search(q, k, path=2)
e1.record()
torch.cuda.synchronize()
st = s.last_stats()
flops = 2.0 * batch * rows * dim
print(f"debug={os.environ.get('ARCHI_TC_DEBUG','0')} rows={rows} dim={dim} {storage} k={k} batch={batch}: coarse {min(ms):.3f} ms "
      f"({flops/min(ms)/1e9:.0f} TFLOP/s, {rows*dim*(2 if storage=='bf16' else 4)/min(ms)/1e6:.0f} GB/s) step {e0.elapsed_time(e1)/10:.3f} ms "
      f"grid={st.grid} unverified={st.unverified_queries}")
