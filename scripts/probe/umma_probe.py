"""Drives umma_probe.cu: checks the K-major layout the forward kernel uses and two candidate readings of the MN-major
(transposed) no-swizzle layout for A and B."""
import ctypes as C, os, sys, json
import numpy as np, torch
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libumma_probe.so"))
dev = torch.device("cuda:0")
M, N, K = 128, 208, 64
rng = np.random.default_rng(0)
A = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).bfloat16()
B = torch.from_numpy(rng.standard_normal((N, K)).astype(np.float32)).bfloat16()
ref = (A.float() @ B.float().T).numpy()

def kmajor(X):
    """(rows R, K): element (r, k) at (k/8)*LBO + (r/8)*128 + (r%8)*16 + (k%8)*2 bytes; LBO = (R/8)*128."""
    R, Kk = X.shape
    img = torch.zeros(R * Kk, dtype=torch.bfloat16)
    r = torch.arange(R).view(-1, 1); k = torch.arange(Kk).view(1, -1)
    off = (k // 8) * (R // 8) * 64 + (r // 8) * 64 + (r % 8) * 8 + (k % 8)
    img[off.reshape(-1)] = X.reshape(-1)
    return img, dict(lbo=(R // 8) * 128, sbo=128, kadv=2 * (R // 8) * 128)

def mnmajor(X):
    """(rows R = MN index, K): element (mn, k) at (k/8)*G + (mn/8)*128 + (k%8)*16 + (mn%8)*2 bytes; G = (R/8)*128:
    a block of 8 K-rows holds, chunk after chunk, 8 MN values (16 B) for each of its 8 rows."""
    R, Kk = X.shape
    img = torch.zeros(R * Kk, dtype=torch.bfloat16)
    r = torch.arange(R).view(-1, 1); k = torch.arange(Kk).view(1, -1)
    G = (R // 8) * 64
    off = (k // 8) * G + (r // 8) * 64 + (k % 8) * 8 + (r % 8)
    img[off.reshape(-1)] = X.reshape(-1)
    return img, G * 2

def idesc(n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((128 >> 4) << 24)

def run(a_img, b_img, da, db, idc, reps=64, n=None):
    n = n or N
    a = a_img.view(torch.uint8).to(dev); b = b_img.view(torch.uint8).to(dev)
    out = torch.full((M, n), float("nan"), device=dev)
    cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    rc = lib.umma_probe(C.c_void_p(a.data_ptr()), a.numel(), C.c_void_p(b.data_ptr()), b.numel(), n, K // 16,
                        da["lbo"], da["sbo"], da["kadv"], db["lbo"], db["sbo"], db["kadv"], idc, C.c_void_p(out.data_ptr()), reps,
                        C.c_void_p(cyc.data_ptr()))
    if rc != 0:
        return f"cuda error {rc}"
    err = np.abs(out.cpu().numpy() - ref[:, :n]).max()
    return {"err": float(err), "cycles_per_mma": int(cyc.item()) / (reps * (K // 16))}

ak, dak = kmajor(A); bk, dbk = kmajor(B)
am, GA = mnmajor(A); bm, GB = mnmajor(B)
cands = {"lbo=G,sbo=128": lambda G: dict(lbo=G, sbo=128, kadv=2 * G), "lbo=128,sbo=G": lambda G: dict(lbo=128, sbo=G, kadv=2 * G)}
good = lambda G: dict(lbo=G, sbo=128, kadv=2 * G)
tests = [("SS  A K-major, B K-major, N=208", lambda: run(ak, bk, dak, dbk, idesc(N, 0, 0))),
         ("SS  A MN-major, B MN-major, N=208", lambda: run(am, bm, good(GA), good(GB), idesc(N, 1, 1))),
         ("SS  A K-major, B K-major, N=16", lambda: run(ak, bk, dak, dbk, idesc(16, 0, 0), n=16)),
         ("SS  A MN-major, B MN-major, N=16", lambda: run(am, bm, good(GA), good(GB), idesc(16, 1, 1), n=16)),
         ("SS  A MN-major, B MN-major, N=208, 1 rep", lambda: run(am, bm, good(GA), good(GB), idesc(N, 1, 1), reps=1))]
if len(sys.argv) > 1:
    name, fn = tests[int(sys.argv[1])]
    print(json.dumps({name: fn()}))
else:
    import subprocess
    for i in range(len(tests)):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True)
        print(r.stdout.strip() or f"{tests[i][0]}: CRASH {r.stderr.strip().splitlines()[-1][:150] if r.stderr.strip() else ''}", flush=True)
