import faulthandler, sys, os
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.getcwd())
import numpy as np
from archi_b200.store import NativeStore
print("create", flush=True)
s = NativeStore(8)
print("append", flush=True)
s.append(np.eye(8, dtype=np.float32))
print("appended", flush=True)
print(s.search(np.eye(8, dtype=np.float32)[:1], 3))
