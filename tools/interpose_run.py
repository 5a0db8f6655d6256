"""Interpose mode check (INTEGRATION.md mode 2), own process: the REAL libmagma here is the compiled reference
(oracle/_ref/libmagma_ref.so, loaded RTLD_GLOBAL). It creates the queue, allocates and fills the device arrays; the
factorisation is then done by libmagma_b200_interpose.so's magma_dgetrf_batched / magma_dgesv_batched on the reference's
opaque queue (stream and device read through the reference's own accessors).
   python tools/interpose_run.py n batch nrhs out.npz"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import oracle

n, batch, nrhs, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmagma_ref.so"), mode=C.RTLD_GLOBAL)
ours = C.CDLL(os.path.join(ROOT, "magma_b200", "lib", "libmagma_b200_interpose.so"), mode=C.RTLD_LOCAL)
vp, i32 = C.c_void_p, C.c_int
assert ref.magma_init() == 0
torch.cuda.set_device(0)
q = vp()
ref.magma_queue_create_internal.argtypes = [i32, vp, C.c_char_p, C.c_char_p, i32]
ref.magma_queue_create_internal(0, C.byref(q), b"f", b"f", 0)     # the reference's queue object
ref.magma_dset_pointer.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
ref.magma_iset_pointer.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
ref.magma_queue_sync_internal.argtypes = [vp, C.c_char_p, C.c_char_p, i32]
ours.magma_dgetrf_batched.argtypes = [i32, i32, vp, i32, vp, vp, i32, vp]
ours.magma_dgesv_batched.argtypes = [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]
dev = torch.device("cuda", 0)
A0, seed = oracle.random_batch(batch, n, n)
A = torch.from_numpy(A0).to(dev)
ipiv = torch.zeros((batch, n), dtype=torch.int32, device=dev)
info = torch.zeros(batch, dtype=torch.int32, device=dev)
pA = torch.zeros(batch, dtype=torch.int64, device=dev)
pP = torch.zeros(batch, dtype=torch.int64, device=dev)
ref.magma_dset_pointer(pA.data_ptr(), A.data_ptr(), n, 0, 0, n * n, batch, q)
ref.magma_iset_pointer(pP.data_ptr(), ipiv.data_ptr(), 1, 0, 0, n, batch, q)
res = {}
if nrhs:
    B0, _ = oracle.random_batch(batch, n, nrhs, iseed=seed)
    B = torch.from_numpy(B0).to(dev)
    pB = torch.zeros(batch, dtype=torch.int64, device=dev)
    ref.magma_dset_pointer(pB.data_ptr(), B.data_ptr(), n, 0, 0, n * nrhs, batch, q)
    rc = ours.magma_dgesv_batched(n, nrhs, pA.data_ptr(), n, pP.data_ptr(), pB.data_ptr(), n, info.data_ptr(), batch, q)
else:
    rc = ours.magma_dgetrf_batched(n, n, pA.data_ptr(), n, pP.data_ptr(), info.data_ptr(), batch, q)
assert rc == 0, rc
ref.magma_queue_sync_internal(q, b"f", b"f", 0)   # the REFERENCE's sync must cover our kernels: same stream
res.update(LU=A.cpu().numpy(), ipiv=ipiv.cpu().numpy(), info=info.cpu().numpy())
if nrhs:
    res["X"] = B.cpu().numpy()
np.savez(out, **res)
print("interpose ok")
