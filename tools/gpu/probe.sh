df -h /dev/shm /tmp | cat
nproc; free -g | head -2
python - <<'PY'
import numpy as np, ctypes as C, os, sys, time
sys.path.insert(0, '.')
import coati_b200
lib = coati_b200.load_library()
ctx = coati_b200.Context(0)
path = "/dev/shm/coati_probe"
n = 3 << 30
with open(path, "wb") as f: f.truncate(n)
m = np.memmap(path, dtype=np.uint8, mode="r+")
t=time.time(); rc = lib.coati_gpu_host_register(C.c_void_p(m.ctypes.data), n); print("register shm rc", rc, time.time()-t)
if rc == 0: print("unregister", lib.coati_gpu_host_unregister(C.c_void_p(m.ctypes.data)))
del m; os.unlink(path)
PY
nvidia-smi topo -m | head -20
