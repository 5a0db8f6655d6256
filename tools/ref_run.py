"""Run the compiled reference GPU path (oracle/_ref/libmagma_ref.so: MAGMA 2.10.0 kernels + cuBLAS) on the oracle's
dlarnv inputs and save its output. Own process: the reference exports the same symbol names as libmagma_b200.so.
   python tools/ref_run.py n batch nrhs out.npz|- [time_reps] [tile] [magma|cublas]
nrhs = 0: magma_dgetrf_batched; nrhs > 0: magma_dgesv_batched. With time_reps, also prints a JSON line with the best
wall-clock time around a queue sync (how the reference testers time, testing/testing_zgetrf_batched.cpp:203-206).
"tile": inputs are 256 dlarnv matrices repeated (timing runs at full BASELINE sizes). "cublas": cublasDgetrfBatched
(the testers' other comparison row, testing/testing_zgetrf_batched.cpp:233-249) instead of the MAGMA driver."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import oracle  # dlarnv stream only

n, batch, nrhs, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tile = "tile" in sys.argv[6:]
impl = "cublas" if "cublas" in sys.argv[6:] else "magma"
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmagma_ref.so"), mode=C.RTLD_LOCAL)
vp, i32 = C.c_void_p, C.c_int
L.magma_init.restype = i32
assert L.magma_init() == 0
torch.cuda.set_device(0)
q = vp()
L.magma_queue_create_internal.argtypes = [i32, vp, C.c_char_p, C.c_char_p, i32]
L.magma_queue_create_internal(0, C.byref(q), b"f", b"f", 0)
L.magma_dset_pointer.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
L.magma_iset_pointer.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
L.magma_dgetrf_batched.argtypes = [i32, i32, vp, i32, vp, vp, i32, vp]
L.magma_dgetrf_batched.restype = i32
L.magma_dgesv_batched.argtypes = [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp]
L.magma_dgesv_batched.restype = i32
L.magma_queue_sync_internal.argtypes = [vp, C.c_char_p, C.c_char_p, i32]
dev = torch.device("cuda", 0)
gen = min(batch, 256) if tile else batch
A0, seed = oracle.random_batch(gen, n, n)
A0d = torch.from_numpy(A0).to(dev)
if gen < batch:
    A0d = A0d.repeat((batch + gen - 1) // gen, 1, 1)[:batch].contiguous()
A = A0d.clone()
ipiv = torch.zeros((batch, n), dtype=torch.int32, device=dev)
info = torch.zeros(batch, dtype=torch.int32, device=dev)
pA = torch.zeros(batch, dtype=torch.int64, device=dev)
pP = torch.zeros(batch, dtype=torch.int64, device=dev)
L.magma_dset_pointer(pA.data_ptr(), A.data_ptr(), n, 0, 0, n * n, batch, q)
L.magma_iset_pointer(pP.data_ptr(), ipiv.data_ptr(), 1, 0, 0, n, batch, q)
B = pB = None
if nrhs:
    B0, _ = oracle.random_batch(gen, n, nrhs, iseed=seed)
    B0d = torch.from_numpy(B0).to(dev)
    if gen < batch:
        B0d = B0d.repeat((batch + gen - 1) // gen, 1, 1)[:batch].contiguous()
    B = B0d.clone()
    pB = torch.zeros(batch, dtype=torch.int64, device=dev)
    L.magma_dset_pointer(pB.data_ptr(), B.data_ptr(), n, 0, 0, n * nrhs, batch, q)
L.magma_queue_sync_internal(q, b"f", b"f", 0)


cub = handle = None
if impl == "cublas":
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cublas", "lib", "libcublas.so*")) + \
        glob.glob("/usr/local/cuda/lib64/libcublas.so*")
    cub = C.CDLL(sorted(cands)[0])
    handle = vp()
    assert cub.cublasCreate_v2(C.byref(handle)) == 0
    cub.cublasDgetrfBatched.argtypes = [vp, i32, vp, i32, vp, vp, i32]


def call():
    if impl == "cublas":
        return cub.cublasDgetrfBatched(handle, n, pA.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), batch)
    if nrhs:
        return L.magma_dgesv_batched(n, nrhs, pA.data_ptr(), n, pP.data_ptr(), pB.data_ptr(), n, info.data_ptr(), batch, q)
    return L.magma_dgetrf_batched(n, n, pA.data_ptr(), n, pP.data_ptr(), info.data_ptr(), batch, q)


rc = call()
L.magma_queue_sync_internal(q, b"f", b"f", 0)
torch.cuda.synchronize()
assert rc == 0, rc
if out != "-":
    res = {"LU": A.cpu().numpy(), "ipiv": ipiv.cpu().numpy(), "info": info.cpu().numpy()}
    if nrhs:
        res["X"] = B.cpu().numpy()
    np.savez(out, **res)
if reps:
    ts = []
    for _ in range(reps):
        A.copy_(A0d)
        if nrhs:
            B.copy_(B0d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        L.magma_queue_sync_internal(q, b"f", b"f", 0)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"n": n, "batch": batch, "nrhs": nrhs, "impl": impl, "ms_best": min(ts)}), flush=True)
