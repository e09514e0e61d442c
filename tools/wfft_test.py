import sys, numpy as np, torch, ctypes
cuda = ctypes.CDLL("libcuda.so.1")
torch.cuda.init(); torch.zeros(1, device="cuda")
mod = ctypes.c_void_p(); fn = ctypes.c_void_p()
assert cuda.cuModuleLoad(ctypes.byref(mod), b"tools/wfft_test.cubin") == 0
assert cuda.cuModuleGetFunction(ctypes.byref(fn), mod, b"wfft_test") == 0
rng = np.random.default_rng(0)
x = (rng.standard_normal(256) + 1j * rng.standard_normal(256)).astype(np.complex64)
d_in = torch.from_numpy(x).cuda(); d_f = torch.zeros(256, dtype=torch.complex64, device="cuda"); d_b = torch.zeros_like(d_f)
args = [ctypes.c_void_p(t.data_ptr()) for t in (d_in, d_f, d_b)]
argv = (ctypes.c_void_p * 3)(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
rc = cuda.cuLaunchKernel(fn, 1, 1, 1, 32, 1, 1, 0, None, argv, None)
torch.cuda.synchronize()
X = np.fft.fft(x.astype(np.complex128))
print("launch rc", rc, "forward rel err", np.abs(d_f.cpu().numpy() - X).max() / np.abs(X).max(),
      "round trip rel err", np.abs(d_b.cpu().numpy() / 256 - x).max() / np.abs(x).max())
