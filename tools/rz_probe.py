"""Probe: how do the B200 tensor cores round their float32 accumulator?

Inputs are exactly representable in TF32 and positive, so every product is exact and any deviation from
the float64 result is accumulation rounding.  Round-to-nearest accumulation has zero mean error; the
tensor-core result shows a negative mean that grows linearly with K (truncation on every MMA step).
Output of one run on a B200 is kept in profiles/r1_tensor_accum_rounding.txt; DESIGN.md section 4.1 uses it
to explain the residual difference between the tcgen05 path and the float32 SIMT path on ill-conditioned
flows.  Library GEMM (cuBLAS through torch) on purpose: it shows the hardware behaviour, independent of
this repo's kernels.
"""
import torch


def trunc(x, bits):
    return (x.view(torch.int32) & ~((1 << (23 - bits)) - 1)).view(torch.float32)


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for K in (8, 32, 128, 512, 2048):
        A = trunc(torch.rand(4096, K, device=dev) + 0.5, 10)
        B = trunc(torch.rand(K, 4096, device=dev) + 0.5, 10)
        ref = A.double() @ B.double()
        torch.backends.cuda.matmul.allow_tf32 = True
        C = (A @ B).double()
        torch.backends.cuda.matmul.allow_tf32 = False
        C32 = (A @ B).double()
        ulp = 2.0 ** (torch.floor(torch.log2(ref)) - 23)
        e = (C - ref) / ulp
        e32 = (C32 - ref) / ulp
        print("K=%5d  tensor core (tf32): mean %8.3f ulp, rms %8.3f ulp | float32 FMA: mean %6.3f ulp, rms %6.3f ulp"
              % (K, e.mean().item(), e.pow(2).mean().sqrt().item(), e32.mean().item(), e32.pow(2).mean().sqrt().item()))


if __name__ == "__main__":
    main()
