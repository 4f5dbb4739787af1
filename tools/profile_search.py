"""Small driver for ncu: builds a synthetic store and runs a few searches.
usage: python tools/profile_search.py [rows] [dim] [storage] [k] [batch] [iters] [path]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from archi_b200.store import NativeStore

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
storage = sys.argv[3] if len(sys.argv) > 3 else "f32"
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 1
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
path = int(sys.argv[7]) if len(sys.argv) > 7 else 0

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
s = NativeStore(dim, "cosine", storage, capacity_rows=rows)
for st in range(0, rows, 262144):
    m = min(262144, rows - st)
    x = torch.randn((m, dim), generator=g, device=dev)
    s.append(x / x.norm(dim=1, keepdim=True))
q = torch.randn((batch, dim), generator=g, device=dev)
q = q / q.norm(dim=1, keepdim=True)
for _ in range(iters):
    sc, ids = s.search(q, k, path=path)
torch.cuda.synchronize()
print("ok", sc[0, :3].tolist(), ids[0, :3].tolist())
