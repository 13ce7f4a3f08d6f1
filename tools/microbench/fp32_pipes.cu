// FP32 issue / pipe micro-benchmarks for the NLM distance kernel (sb_nlm3.cu): which instruction
// forms reach the 128 FP32 lanes per SM per clock on sm_100a, and whether the packed
// FFMA2 / FADD2 forms (fma.rn.f32x2, add.rn.f32x2) halve the issue slots of the
// (a - b)^2 accumulate that the row terms are made of.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 2048
#define NV 16

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float v[NV], s[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {      // opaque values: loaded, so nothing folds into an immediate
    v[i] = out[(threadIdx.x + 32 * i) & 1023];
    s[i] = out[(threadIdx.x + 32 * i + 512) & 1023];
  }
  float ra = a + (float)threadIdx.x * 1e-9f, rb = b + (float)threadIdx.x * 1e-9f;   // registers, not constants
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {            // FFMA with constant-bank operands: v = v*c[a] + c[b]
#pragma unroll
      for (int i = 0; i < NV; i++) v[i] = fmaf(v[i], a, b);
    } else if (MODE == 1) {     // FFMA, three register operands
#pragma unroll
      for (int i = 0; i < NV; i++) v[i] = fmaf(v[i], ra, rb);
    } else if (MODE == 2) {     // FADD reg,reg
#pragma unroll
      for (int i = 0; i < NV; i++) v[i] = v[i] + ra;
    } else if (MODE == 3) {     // the row-term pattern: d = t - s; r = fma(d, d, r)
#pragma unroll
      for (int i = 0; i < NV; i++) { const float d = v[i] - s[i]; s[i] = fmaf(d, d, s[i]); }   // d depends on the accumulator: nothing to hoist
    } else if (MODE == 4) {     // packed: FFMA2 three-register
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        float2 x = make_float2(v[i], v[i + 1]);
        x = __ffma2_rn(x, make_float2(ra, rb), make_float2(rb, ra));
        v[i] = x.x; v[i + 1] = x.y;
      }
    } else if (MODE == 5) {     // packed row-term pattern: d2 = t2 - s2 (FADD2 with negated operand), r2 = fma2(d2, d2, r2)
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const float2 d = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(-s[i], -s[i + 1]));
        float2 x = __ffma2_rn(d, d, make_float2(s[i], s[i + 1]));
        s[i] = x.x; s[i + 1] = x.y;
      }
    } else if (MODE == 6) {     // FMNMX (alu pipe) interleaved with FFMA: do they co-issue?
#pragma unroll
      for (int i = 0; i < NV; i += 2) { v[i] = fmaf(v[i], ra, rb); v[i + 1] = fminf(v[i + 1], s[i]); s[i] += 1.0f; }
    } else if (MODE == 8) {     // FFMA2 + FMNMX 1:1 -- does the ALU instruction issue under the 2-cycle FFMA2?
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        float2 x = __ffma2_rn(make_float2(v[i], v[i + 1]), make_float2(ra, rb), make_float2(rb, ra));
        v[i] = x.x; v[i + 1] = x.y;
        s[i] = fminf(s[i], x.x + 0.f * s[i + 1]) ; s[i + 1] = fmaxf(s[i + 1], ra);
      }
    } else if (MODE == 9) {     // FFMA + FMNMX 1:1
#pragma unroll
      for (int i = 0; i < NV; i++) { v[i] = fmaf(v[i], ra, rb); s[i] = fminf(s[i], v[(i + 3) % NV]); }
    } else if (MODE == 10) {    // FFMA2 + 2 FMNMX per FFMA2 (same useful work as mode 9)
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        float2 x = __ffma2_rn(make_float2(v[i], v[i + 1]), make_float2(ra, rb), make_float2(rb, ra));
        v[i] = x.x; v[i + 1] = x.y;
        s[i] = fminf(s[i], v[(i + 3) % NV]); s[i + 1] = fminf(s[i + 1], v[(i + 4) % NV]);
      }
    } else if (MODE == 7) {     // FADD2 alone
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        float2 x = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(ra, rb));
        v[i] = x.x; v[i + 1] = x.y;
      }
    }
  }
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) t += v[i] + s[i];
  if (t == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
static void run(const char* name, double flop_per_elem, double instr_per_elem) {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double elems = (double)ITERS * NV * 148 * 8 * 256;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr_per_clk_per_sm = elems * instr_per_elem / 32.0 / (best * 1e-3) / 148.0 / (clk * 1e3);
  printf("%-44s %8.3f ms  %7.2f TFLOP/s  %5.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, best,
         elems * flop_per_elem / (best * 1e-3) / 1e12, warp_instr_per_clk_per_sm, clk / 1000);
  cudaFree(d);
}

int main() {
  run<0>("FFMA const operands", 2, 1);
  run<1>("FFMA reg,reg,reg", 2, 1);
  run<2>("FADD reg,reg", 1, 1);
  run<3>("FADD + FFMA(d,d,r)  [row-term pattern]", 3, 2);
  run<4>("FFMA2 reg pairs", 2, 0.5);
  run<5>("FADD2 + FFMA2(d,d,r) [packed row term]", 3, 1);
  run<6>("FFMA + FMNMX + FADD interleaved", 2, 1.5);
  run<7>("FADD2 reg pairs", 1, 0.5);
  run<9>("FFMA + FMNMX 1:1", 2, 2);
  run<10>("FFMA2 + 2 FMNMX (same work as above)", 2, 1.5);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
