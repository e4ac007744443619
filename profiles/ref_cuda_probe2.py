"""REF_CUDA probe 2: which summation order do torch-CUDA's tiny batched matmuls ([B,3,3]@[B,3,3], [B,3,3]@[B,3,4]: cuBLAS)
use, and do CUDA's sinf / cosf (what a kernel gets from libdevice) equal torch.sin / torch.cos on CUDA tensors?
usage: python profiles/ref_cuda_probe2.py"""
import numpy as np, torch

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(3)
N = 200000


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fma(a, b, c):   # fp32 fma through fp64 (product exact; the double rounding of the sum is a ~1e-9 event)
    return f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def candidates(A, Bm):
    a = lambda k: A[:, :, k, None]          # [N,3,1]
    b = lambda k: Bm[:, None, k, :]         # [N,1,M]
    p = [f32(a(k).astype(np.float64) * b(k).astype(np.float64)) for k in range(3)]
    out = {}
    out["mul, add, add (torch-CPU's form)"] = f32(f32(p[0].astype(np.float64) + p[1]).astype(np.float64) + p[2])
    out["fma chain, k ascending"] = fma(a(2), b(2), fma(a(1), b(1), p[0]))
    out["fma chain, k descending"] = fma(a(0), b(0), fma(a(1), b(1), p[2]))
    out["p0 + fma(a1,b1,p2)"] = f32(p[0].astype(np.float64) + fma(a(1), b(1), p[2]))
    return out


def report(tag, A, Bm):
    ref = (torch.from_numpy(A).cuda() @ torch.from_numpy(Bm).cuda()).cpu().numpy()
    cpu = (torch.from_numpy(A) @ torch.from_numpy(Bm)).numpy()
    print(f"{tag}: torch-CUDA vs torch-CPU {np.mean(ref.view(np.int32) == cpu.view(np.int32)) * 100:.3f} % bit-identical")
    for name, c in candidates(A, Bm).items():
        print(f"    torch-CUDA == {name:36s}: {np.mean(ref.view(np.int32) == c.view(np.int32)) * 100:.4f} %")


A = torch.randn(N, 3, 3, generator=g).numpy()
B3 = torch.randn(N, 3, 3, generator=g).numpy()
B4 = torch.randn(N, 3, 4, generator=g).numpy()
report("[N,3,3]@[N,3,3] random", A, B3)
report("[N,3,3]@[N,3,4] random", A, B4)
for n in (4, 64, 256):
    report(f"[{n},3,3]@[{n},3,4] random (batch {n})", A[:n].copy(), B4[:n].copy())
# trigonometry: CUDA libdevice through a raw kernel vs torch on CUDA
x = torch.cat([torch.randn(N, generator=g) * s for s in (0.01, 0.3, 3.0, 100.0)]).cuda()
src = r'''
extern "C" __global__ void trig(const float* x, float* s, float* c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); }
}'''
try:
    from torch.utils.cpp_extension import load_inline   # noqa: F401  (not used: NVRTC through cuda-python is lighter)
    from cuda import nvrtc, cuda as cu
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"trig.cu", 0, [], [])
    opts = [b"--gpu-architecture=compute_100a", b"--fmad=false"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    err, sz = nvrtc.nvrtcGetPTXSize(prog)
    ptx = b" " * sz
    nvrtc.nvrtcGetPTX(prog, ptx)
    err, mod = cu.cuModuleLoadData(np.char.array(ptx).ctypes.data)
    err, fn = cu.cuModuleGetFunction(mod, b"trig")
    s = torch.empty_like(x); c = torch.empty_like(x)
    n = np.array([x.numel()], np.int32)
    args = [np.array([x.data_ptr()], np.uint64), np.array([s.data_ptr()], np.uint64), np.array([c.data_ptr()], np.uint64), n]
    argp = np.array([a.ctypes.data for a in args], np.uint64)
    torch.cuda.synchronize()
    err, = cu.cuLaunchKernel(fn, (x.numel() + 255) // 256, 1, 1, 256, 1, 1, 0, torch.cuda.current_stream().cuda_stream, argp.ctypes.data, 0)
    torch.cuda.synchronize()
    eq = lambda a, b: float((a.view(torch.int32) == b.view(torch.int32)).float().mean()) * 100
    print(f"sinf (libdevice, nvrtc) == torch.sin on CUDA: {eq(s, torch.sin(x)):.4f} %   cosf == torch.cos: {eq(c, torch.cos(x)):.4f} %")
    print(f"torch.sin CUDA == torch.sin CPU: {eq(torch.sin(x).cpu(), torch.sin(x.cpu())):.4f} %   cos: {eq(torch.cos(x).cpu(), torch.cos(x.cpu())):.4f} %")
except Exception as e:   # the matmul part is the point of this probe
    print("trig probe skipped:", type(e).__name__, e)
