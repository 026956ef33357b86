// tools/ubench/pipes.cu -- instruction-issue micro-benchmark for the integer pipes of sm_100a.
// Measures thread-instructions per clock per SM for the SASS forms the mod-P butterflies are made of,
// alone and mixed, so that kernel design can be budgeted per pipe (alu vs fma) instead of guessed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITER 4096
#define ACC 8

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t mulc) {
    uint32_t a[ACC], b[ACC];
    uint64_t w[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + threadIdx.x; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    uint32_t m = mulc;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if constexpr (KIND == 0) {          // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
            } else if constexpr (KIND == 1) {   // 64-bit add: IADD3 + IADD3.X
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(seed));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(m));
            } else if constexpr (KIND == 2) {   // IMAD lo
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(a[i]));
            } else if constexpr (KIND == 3) {   // IMAD.WIDE.U32
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(a[i]));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(m));
            } else if constexpr (KIND == 4) {   // IMAD.HI.U32
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(a[i]));
            } else if constexpr (KIND == 5) {   // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[i]) : "r"(m), "r"(a[i]));
            } else if constexpr (KIND == 6) {   // SHF funnel
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(b[i]) : "r"(a[i]));
            } else if constexpr (KIND == 7) {   // mix 1:1 IADD3 : IMAD
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(seed));
            } else if constexpr (KIND == 8) {   // mix 1:1 IADD3 : IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(seed));
            } else if constexpr (KIND == 9) {   // mix 2:1 IADD3 : IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(seed));
            } else if constexpr (KIND == 10) {  // IMAD.WIDE with carry out + carry in (mad.cc chain) : 64x32 -> 96 MAC
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(seed));
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(m));
            } else if constexpr (KIND == 11) {  // 64-bit mul.hi + mul.lo
                uint64_t hi, lo;
                asm volatile("mul.hi.u64 %0, %2, %3;\n\tmul.lo.u64 %1, %2, %3;" : "=l"(hi), "=l"(lo) : "l"(w[i]), "l"((uint64_t)m << 13 | seed));
                w[i] = hi ^ lo;
            } else if constexpr (KIND == 12) {  // LEA (shift-add)
                asm volatile("{.reg .u32 t; shl.b32 t, %1, 5; add.u32 %0, %0, t;}" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("{.reg .u32 t; shl.b32 t, %1, 9; add.u32 %0, %0, t;}" : "+r"(b[i]) : "r"(a[i]));
            } else if constexpr (KIND == 13) {  // 3-input add
                asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(a[i]) : "r"(b[i]), "r"(m));
                asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(b[i]) : "r"(a[i]), "r"(m));
            } else if constexpr (KIND == 14) {  // DFMA
                double x = __longlong_as_double(w[i]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(1.0000001), "d"(0.5));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(0.9999999), "d"(0.25));
                w[i] = __double_as_longlong(x);
            } else if constexpr (KIND == 15) {  // mix IADD3 : IMAD : 1:1 with 64-bit add chains (2 IADD3-class) + 2 IMAD
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(((uint32_t*)&w[i])[0]) : "r"(m), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(((uint32_t*)&w[i])[1]) : "r"(m), "r"(seed));
            } else if constexpr (KIND == 16) {  // IMAD.SHL-like: mul by power of two constant + add (imm form)
                asm volatile("mad.lo.u32 %0, %0, 256, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, 65536, %1;" : "+r"(b[i]) : "r"(a[i]));
            } else if constexpr (KIND == 17) {  // mix 1:1:1  IADD3 : IMAD : IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(seed));
            } else if constexpr (KIND == 18) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(b[i]) : "r"(a[i]));
            } else if constexpr (KIND == 19) {  // mix 3:1 IADD3 : IMAD.WIDE
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(m));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(m), "r"(seed));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

struct Kind { const char* name; int per_iter; };   // PTX-level instructions per inner (i) iteration
static const Kind kinds[] = {
    {"IADD3 (32-bit add)", 2}, {"IADD3+IADD3.X (64-bit add)", 4}, {"IMAD lo", 2}, {"IMAD.WIDE.U32", 2}, {"IMAD.HI.U32", 2},
    {"LOP3", 2}, {"SHF funnel", 2}, {"mix IADD3:IMAD 1:1", 2}, {"mix IADD3:IMAD.WIDE 1:1", 2}, {"mix IADD3:IMAD.WIDE 2:1", 3},
    {"mad.lo.cc+madc.hi", 4}, {"mul.hi.u64+mul.lo.u64 (+xor)", 2}, {"shl+add (LEA)", 2}, {"add3", 2}, {"DFMA", 2},
    {"64-bit add + 2 IMAD", 4}, {"IMAD by 2^k imm + add", 2}, {"mix IADD3:IMAD:IMAD.WIDE", 3}, {"PRMT", 2}, {"mix IADD3:IMAD.WIDE 3:1", 4}};

template <int KIND>
void run(uint32_t* out, int sms, double ghz) {
    const int blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<blocks, 256>>>(out, 12345u, 77u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND><<<blocks, 256>>>(out, 12345u, 77u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * ITER * ACC * kinds[KIND].per_iter;
    double per_clk_sm = ops / (ms * 1e-3) / (ghz * 1e9) / sms;
    printf("%-34s %8.3f ms  %7.1f thread-instr/clk/SM (PTX-level count, at %.3f GHz)\n", kinds[KIND].name, ms, per_clk_sm, ghz);
}
template <int... K> void run_all(uint32_t* out, int sms, double ghz, std::integer_sequence<int, K...>) { (run<K>(out, sms, ghz), ...); }

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    printf("%s: %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    uint32_t* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    run_all(out, p.multiProcessorCount, ghz, std::make_integer_sequence<int, 20>{});
    return 0;
}
