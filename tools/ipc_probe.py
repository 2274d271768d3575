"""Probe: can one process per GPU map a peer's cudaMalloc buffer through CUDA IPC here? (torchrun, 2 ranks)"""
import ctypes, os, sys
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
world = dist.get_world_size()
rt = ctypes.CDLL("libcudart.so.12")
ptr = ctypes.c_void_p()
assert rt.cudaSetDevice(lr) == 0
assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(1 << 20)) == 0
val = (ctypes.c_int * 4)(rank + 100, 1, 2, 3)
assert rt.cudaMemcpy(ptr, val, ctypes.c_size_t(16), 1) == 0
handle = (ctypes.c_ubyte * 64)()
rc = rt.cudaIpcGetMemHandle(handle, ptr)
print(rank, "get handle rc", rc, flush=True)
handles = [None] * world
dist.all_gather_object(handles, bytes(handle))
peer = (rank + 1) % world
ph = (ctypes.c_ubyte * 64).from_buffer_copy(handles[peer])
pptr = ctypes.c_void_p()
# cudaIpcOpenMemHandle takes the 64-byte struct BY VALUE
class H(ctypes.Structure):
    _fields_ = [("r", ctypes.c_ubyte * 64)]
hh = H(); ctypes.memmove(ctypes.byref(hh), ph, 64)
rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), H, ctypes.c_uint]
rc = rt.cudaIpcOpenMemHandle(ctypes.byref(pptr), hh, 1)
print(rank, "open handle rc", rc, hex(pptr.value or 0), flush=True)
out = (ctypes.c_int * 4)()
rc = rt.cudaMemcpy(out, pptr, ctypes.c_size_t(16), 2)
print(rank, "peer read rc", rc, list(out), flush=True)
dist.barrier()
dist.destroy_process_group()
