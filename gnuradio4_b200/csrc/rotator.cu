// Complex mixer: Rotator<std::complex<float>> (blocks/math/include/gnuradio-4.0/math/Rotator.hpp:40-61).
//
// The reference advances ONE float phase accumulator per sample,
//     phase += dphi;  if (phase > 2pi_f) phase -= 2pi_f;  else if (phase < 0) phase += 2pi_f;
// so sample i is rotated by the i-fold iterate of a float -> float map. Rounding makes that iterate drift away from
// i*dphi (the bias per step is systematic, not random), hence a closed form cannot reproduce the reference; the
// recurrence itself has to be replayed. It is input independent, which is what makes it parallel:
//
//  * Every wrap lands on a coarse grid: for dphi > 0 the value after `phase -= 2pi_f` is an exact multiple of 2^-21 in
//    [0, dphi]; for dphi < 0 the value after `phase += 2pi_f` is 2pi_f minus an exact multiple of 2^-22 (|dphi| <= pi).
//    So "phase right after a wrap" takes at most |dphi| * 2^22 + 2 distinct values: the landing states.
//  * base table  T0[k] = (landing state reached by the NEXT wrap, number of samples in between), built by simulating
//    one revolution from every landing state (one thread each);
//  * T_j = T_{j-1} o T_{j-1}: 2^j revolutions per lookup (binary lifting);
//  * the phase in front of any sample index m is then: shared prefix up to the first landing, ~log2(m) table lookups,
//    and fewer than one revolution of plain replay. One thread does this per 4096-sample tile and then replays its tile,
//    dropping a checkpoint every 16 samples; the main kernel replays 16 steps per thread from those checkpoints.
// Every float operation of the reference is executed, in the reference's order, by some thread => bit-identical phases.
// cos/sin of the phase are the C library's sinf / cosf restated operation by operation (sincos_core.cuh, FP64 pipe), the
// product uses std::complex rounding => the OUTPUT is bit-identical to the reference as well (tests/test_gpu_parity.py).
// |dphi| > pi, non-finite dphi or a stalled accumulator (dphi below half an ulp of the phase) use a serial replay.
#include <cmath>
#include <cstring>
#include <vector>

#include <cstdlib>
#include <algorithm>

#include "common.cuh"
#include "rotator_core.cuh"

namespace gr4b200 {
namespace {

__global__ void prefixKernel(Landing l, const float* __restrict__ startPhase, unsigned long long nSamples, Prefix* __restrict__ prefix) { computePrefix(l, *startPhase, nSamples, prefix); }

__global__ void baseTableKernel(Landing l, unsigned long long* __restrict__ table, unsigned long long maxSteps, int* __restrict__ failed) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= l.nStates) {
        return;
    }
    bool gridViolated = false;
    table[k]          = baseTableEntry(l, k, maxSteps, gridViolated);
    if (gridViolated) {
        atomicExch(failed, 1);
    }
}

__global__ void liftTableKernel(const unsigned long long* __restrict__ prev, unsigned long long* __restrict__ next, int nStates) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nStates) {
        next[k] = liftTableEntry(prev, k);
    }
}

// one thread per 512-sample stretch: phase in front of it via the tables, then replay it and keep a checkpoint per run
// (written as 16-byte vectors). Thread nStretches (one past the end) only computes the phase after the last sample and
// stores it as the new state. Many short stretches, not few long ones: the replay is a serial float recurrence, its
// latency is hidden by thread count only.
template<int kCheckpointTile>
__global__ void __launch_bounds__(128) checkpointKernel(Landing l, const float* __restrict__ startPhase, const Prefix* __restrict__ prefix, const unsigned long long* __restrict__ tables, int nLevels, unsigned long long nSamples, float* __restrict__ runPhases, float* __restrict__ endPhase) {
    constexpr int            kRuns      = kCheckpointTile / kRun;
    const unsigned long long nStretches = (nSamples + kCheckpointTile - 1) / kCheckpointTile;
    const unsigned long long stretch    = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (stretch > nStretches) {
        return;
    }
    const Prefix pre = *prefix;
    if (stretch == nStretches) {
        *endPhase = phaseBeforeSample(l, *startPhase, pre, tables, nLevels, nSamples);
        return;
    }
    float                    phase    = phaseBeforeSample(l, *startPhase, pre, tables, nLevels, stretch * kCheckpointTile);
    const unsigned long long firstRun = stretch * kRuns;
    const unsigned long long lastRun  = (nSamples + kRun - 1) / kRun; // checkpoints exist for runs [0, lastRun)
    const float              dphi     = l.dphi;
#pragma unroll 1
    for (int r4 = 0; r4 < kRuns; r4 += 4) {
        float cp[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            cp[r] = phase;
#pragma unroll
            for (int i = 0; i < kRun; ++i) { // stepping past the end of the call is harmless: those phases are not used
                bool wrapped;
                phase = stepPhase(phase, dphi, wrapped);
            }
        }
        if (firstRun + r4 + 4 <= lastRun) { // runPhases comes from cudaMalloc and firstRun + r4 is a multiple of 4
            *reinterpret_cast<float4*>(runPhases + firstRun + r4) = make_float4(cp[0], cp[1], cp[2], cp[3]);
        } else {
            for (int r = 0; r < 4 && firstRun + r4 + r < lastRun; ++r) {
                runPhases[firstRun + r4 + r] = cp[r];
            }
            break;
        }
    }
}

// serial fallback: a single thread replays the call (exact for any dphi). An accumulator that no longer moves -- a zero
// increment once the phase has been wrapped into range (the default Rotator), or an increment below half an ulp of the
// phase -- is detected: the replay stops there, `settled` receives the sample index and the fixed phase, and
// fillSettledKernel writes the remaining checkpoints in parallel. Only |dphi| > pi pays the full serial replay.
struct Settled {
    unsigned long long steps; // samples replayed serially (a multiple of kRun unless the call ended first)
    float              phase; // the phase from there on
};
__global__ void serialCheckpointKernel(float dphi, const float* __restrict__ startPhase, unsigned long long nSamples, float* __restrict__ runPhases, float* __restrict__ endPhase, Settled* __restrict__ settled) {
    float              phase = *startPhase;
    unsigned long long i     = 0;
    for (; i < nSamples; ++i) {
        if (i % kRun == 0) {
            runPhases[i / kRun] = phase;
            bool        wrapped;
            const float next = stepPhase(phase, dphi, wrapped);
            if (next == phase) { // fixed point of the recurrence: every later phase is this one
                break;
            }
        }
        bool wrapped;
        phase = stepPhase(phase, dphi, wrapped);
    }
    settled->steps = i < nSamples ? i : (nSamples + kRun - 1) / kRun * kRun; // ran to the end: nothing left to fill
    settled->phase = phase;
    *endPhase      = phase;
}
__global__ void fillSettledKernel(const Settled* __restrict__ settled, unsigned long long nSamples, float* __restrict__ runPhases) {
    const unsigned long long firstRun = settled->steps / kRun, lastRun = (nSamples + kRun - 1) / kRun;
    for (unsigned long long r = firstRun + static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < lastRun; r += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        runPhases[r] = settled->phase;
    }
}

// main kernel: CTA = one tile of 4096 samples, 256 threads. Thread t replays runs t and t + 256 (8 steps each, two
// independent chains) into shared memory, then the CTA rotates the tile with coalesced 16-byte accesses.
// Checkpoints from the plan's phase cycle (PhaseCycle below): the phase in front of sample m of the plan's life is
// cycle[m] for m < size, cycle[mu + (m - mu) mod lambda] beyond; the table is stored as eight residue tables
// (entry q at [q mod 8][q / 8]) so that the checkpoints of consecutive runs of 8 samples are consecutive floats.
constexpr int kGatherPerThread = 16;
__global__ void __launch_bounds__(256) gatherCheckpointKernel(const float* __restrict__ table, unsigned long long stride, unsigned long long mu, unsigned long long lambda, unsigned long long size, unsigned long long position, unsigned long long nRuns, float* __restrict__ runPhases) {
    // a CTA owns 256 x 16 consecutive runs; thread t takes runs t, t + 256, ...: neighbouring threads read neighbouring
    // table entries and write neighbouring checkpoints, and a thread pays the 64-bit modulo once
    const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * (256 * kGatherPerThread) + threadIdx.x;
    if (base >= nRuns) {
        return;
    }
    const unsigned long long m = position + kRun * base;
    unsigned long long       q = m < size ? m : mu + (m - mu) % lambda;
#pragma unroll
    for (int i = 0; i < kGatherPerThread; ++i) {
        const unsigned long long run = base + 256ull * i;
        if (run < nRuns) {
            runPhases[run] = table[(q & 7ull) * stride + (q >> 3)];
        }
        q += kRun * 256ull;
        if (q >= size) { // at most a few periods back (a period shorter than the stride takes the modulo)
            q = q - size < lambda ? q - lambda : mu + (q - mu) % lambda;
        }
    }
}

// Sign: +1 / -1 = the sign of a phase increment with 0 < |dphi| <= pi, 0 = anything else. With a known sign, a tile whose
// checkpoints all lie in [0, 2 pi_f] (every phase after the first wrap does) takes the one-sided wrap test and the sin/cos
// without the sin(-0) select (rotator_core.cuh stepPhaseInRange, common.cuh mixerSinCosInRange): same bits, fewer instructions.
template<int Sign>
__global__ void __launch_bounds__(256) rotateKernel(const float2* __restrict__ in, float2* __restrict__ out, unsigned long long nSamples, float dphi, const float* __restrict__ runPhases) {
    __shared__ float sPhase[kRunsPerTile * (kRun + 1)];
    const unsigned long long nTiles = (nSamples + kTile - 1) / kTile;
    const int                t      = threadIdx.x;
    for (unsigned long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const unsigned long long first = tile * kTile;
        const bool               aligned  = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
        const bool               fullTile = aligned && first + kTile <= nSamples;
        float4                   v[kTile / 2 / 256];
        if (fullTile) { // the samples do not depend on the phases: their loads fly while the phases are replayed
            const float4* in4 = reinterpret_cast<const float4*>(in + first);
#pragma unroll
            for (int u = 0; u < kTile / 2 / 256; ++u) {
                v[u] = ldStream4(in4 + u * 256 + t);
            }
        }
        bool inRange = Sign != 0; // CTA-uniform after the barrier below
        {
            constexpr int PerThread = kRunsPerTile / 256;
            float         phase[PerThread];
#pragma unroll
            for (int b = 0; b < PerThread; ++b) {
                const unsigned long long run = first / kRun + t + b * 256;
                phase[b]                     = run * kRun < nSamples ? runPhases[run] : 0.f; // past the end: replayed, never used
                inRange                      = inRange && phase[b] >= 0.f && phase[b] <= kTwoPi;
            }
            inRange = __syncthreads_and(inRange ? 1 : 0) != 0; // also the barrier in front of the writes to sPhase
            if (inRange) {
#pragma unroll
                for (int i = 0; i < kRun; ++i) {
#pragma unroll
                    for (int b = 0; b < PerThread; ++b) {
                        phase[b]                                  = stepPhaseInRange<(Sign > 0)>(phase[b], dphi);
                        sPhase[(t + b * 256) * (kRun + 1) + i] = phase[b];
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < kRun; ++i) {
#pragma unroll
                    for (int b = 0; b < PerThread; ++b) {
                        bool wrapped;
                        phase[b]                                  = stepPhase(phase[b], dphi, wrapped); // Rotator.hpp:52-58: increment first, then use
                        sPhase[(t + b * 256) * (kRun + 1) + i] = phase[b];
                    }
                }
            }
        }
        __syncthreads();
        if (fullTile) {
            // straight-line code for the whole tile: no calls inside the loop (the out-of-line paths -- a phase outside the
            // one-step reduction range, a product with both parts NaN -- would force the compiler to re-materialise every
            // double constant around each call site); a thread that meets either redoes its samples afterwards
            float4* out4 = reinterpret_cast<float4*>(out + first);
            bool    ok   = true;
#pragma unroll
            for (int u = 0; u < kTile / 2 / 256; ++u) {
                const int   s = 2 * (u * 256 + t); // tile-relative index of the first of two samples
                const float p0 = sPhase[(s / kRun) * (kRun + 1) + s % kRun], p1 = sPhase[((s + 1) / kRun) * (kRun + 1) + (s + 1) % kRun];
                float       c0, s0, c1, s1;
                if (inRange) { // phases in [0, 2 pi_f], never -0: no range test, no sin(-0) select
                    mixerSinCosInRange(p0, &s0, &c0);
                    mixerSinCosInRange(p1, &s1, &c1);
                } else {
                    ok = ok && fabsf(p0) < kSinCosSmallLimit && fabsf(p1) < kSinCosSmallLimit;
                    mixerSinCosFast(p0, &s0, &c0);
                    mixerSinCosFast(p1, &s1, &c1);
                }
                const float ax = __fsub_rn(__fmul_rn(v[u].x, c0), __fmul_rn(v[u].y, s0)), ay = __fadd_rn(__fmul_rn(v[u].x, s0), __fmul_rn(v[u].y, c0));
                const float bx = __fsub_rn(__fmul_rn(v[u].z, c1), __fmul_rn(v[u].w, s1)), by = __fadd_rn(__fmul_rn(v[u].z, s1), __fmul_rn(v[u].w, c1));
                ok = ok && !(ax != ax || ay != ay || bx != bx || by != by); // any NaN part (a superset of Annex G's "both"): general form below
                stStream4(out4 + u * 256 + t, make_float4(ax, ay, bx, by));
            }
            if (!ok) { // rare: the general forms (library reduction for far phases, Annex G recovery of the product)
#pragma unroll 1
                for (int u = 0; u < kTile / 2 / 256; ++u) {
                    const int s = 2 * (u * 256 + t);
                    float     c0, s0, c1, s1;
                    mixerSinCos(sPhase[(s / kRun) * (kRun + 1) + s % kRun], &s0, &c0);
                    mixerSinCos(sPhase[((s + 1) / kRun) * (kRun + 1) + (s + 1) % kRun], &s1, &c1);
                    const float2 a = complexMulAnnexG(v[u].x, v[u].y, c0, s0);
                    const float2 b = complexMulAnnexG(v[u].z, v[u].w, c1, s1);
                    stStream4(out4 + u * 256 + t, make_float4(a.x, a.y, b.x, b.y));
                }
            }
        } else {
            for (int s = t; s < kTile && first + s < nSamples; s += 256) {
                float        c, sn;
                mixerSinCos(sPhase[(s / kRun) * (kRun + 1) + s % kRun], &sn, &c);
                const float2 x = in[first + s];
                out[first + s] = complexMulAnnexG(x.x, x.y, c, sn);
            }
        }
    }
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

struct gr4b200_rotator_plan {
    int                 device     = 0;       // the device the plan's memory lives on
    float               dphi       = 0.f;
    float*              phase      = nullptr; // device: accumulated phase (Rotator::_accumulated_phase)
    float*              endPhase   = nullptr; // device scratch
    Settled*            settled    = nullptr; // device scratch of the serial path
    Prefix*             prefix     = nullptr; // device
    int*                failed     = nullptr; // device flag
    unsigned long long* tables     = nullptr; // device [nLevels][nStates]
    int                 nLevels    = 0;
    unsigned long long  coveredSteps = 0;     // tables are valid for calls up to this many samples
    float*              runPhases  = nullptr; // device scratch, one float per kRun samples
    size_t              runCapacity = 0;
    Landing             landing{};
    bool                useTables  = false;
    // The phase recurrence is a map on the 2^32 float patterns, so the phase sequence of a plan is eventually periodic:
    // phase in front of sample m = cycle[m] for m < mu + lambda, periodic with lambda from mu on. The host replays the
    // recurrence once per plan (and per set_phase) until a phase right after a wrap repeats -- at most about
    // 2 pi * 2^22 = 26 M steps, the number of landing states times the steps per revolution -- and the checkpoints of every
    // call are then gathered from that table instead of being looked up and replayed per stretch.
    float               startPhase = 0.f;     // host: phase in front of sample 0 of the table's origin
    unsigned long long  position   = 0;       // samples consumed since that origin
    bool                cycleTried = false;
    bool                cycleValid = false;
    unsigned long long  cycleMu = 0, cycleLambda = 0, cycleStride = 0;
    float*              cycleTable = nullptr; // device: 8 residue tables of cycleStride floats
    unsigned long long  pendingSamples = 0;   // samples of the call being issued (added to position at commit)
};

namespace {

int ensureTables(gr4b200_rotator_plan* plan, unsigned long long nSamples, cudaStream_t stream) {
    if (!plan->useTables || nSamples <= plan->coveredSteps) {
        return GR4B200_OK;
    }
    const int levels = liftingLevels(plan->dphi, nSamples);
    const size_t entries = static_cast<size_t>(levels) * plan->landing.nStates;
    if (plan->tables != nullptr) {
        GR4B200_CUDA_TRY(cudaStreamSynchronize(stream));
        GR4B200_CUDA_TRY(cudaFree(plan->tables));
        plan->tables = nullptr;
    }
    if (cudaMalloc(&plan->tables, entries * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        plan->useTables = false; // not enough memory for the tables: serial replay still gives the exact answer
        return GR4B200_OK;
    }
    const int threads = 128;
    const int blocks  = ceilDiv(plan->landing.nStates, threads);
    GR4B200_CUDA_TRY(cudaMemsetAsync(plan->failed, 0, sizeof(int), stream));
    baseTableKernel<<<blocks, threads, 0, stream>>>(plan->landing, plan->tables, 1ull << 26, plan->failed);
    for (int j = 1; j < levels; ++j) {
        liftTableKernel<<<blocks, threads, 0, stream>>>(plan->tables + static_cast<size_t>(j - 1) * plan->landing.nStates, plan->tables + static_cast<size_t>(j) * plan->landing.nStates, plan->landing.nStates);
    }
    int failed = 0;
    GR4B200_CUDA_TRY(cudaMemcpyAsync(&failed, plan->failed, sizeof(int), cudaMemcpyDeviceToHost, stream));
    GR4B200_CUDA_TRY(cudaStreamSynchronize(stream)); // one-off, at plan warm-up / growth only
    if (failed != 0) {
        plan->useTables = false;
        return GR4B200_OK;
    }
    plan->nLevels      = levels;
    plan->coveredSteps = nSamples;
    return checkLaunch("rotator tables", static_cast<unsigned>(levels));
}

} // namespace

namespace {

// GR4B200_ROTATOR_CYCLE=0: keep the per-call table look-ups (A/B timing and the fallback's test coverage)
bool cycleEnabled() {
    static const bool enabled = [] { const char* e = std::getenv("GR4B200_ROTATOR_CYCLE"); return e == nullptr || e[0] != '0'; }();
    return enabled;
}

// replays the recurrence on the host until it closes (same float operations as the device: add, compare, add)
void buildPhaseCycle(gr4b200_rotator_plan* plan) {
    plan->cycleTried = true;
    plan->cycleValid = false;
    const float dphi = plan->dphi;
    if (!cycleEnabled() || !std::isfinite(dphi) || !std::isfinite(plan->startPhase)) {
        return;
    }
    std::vector<float> cycle;
    unsigned long long mu = 0, lambda = 0;
    if (!findPhaseCycle(dphi, plan->startPhase, cycle, mu, lambda)) {
        return;
    }
    const unsigned long long size   = mu + lambda;
    const unsigned long long stride = (size + 7) / 8;
    std::vector<float>       residues(8 * stride, 0.f);
    for (unsigned long long q = 0; q < size; ++q) {
        residues[(q & 7ull) * stride + (q >> 3)] = cycle[q];
    }
    if (plan->cycleTable != nullptr) {
        cudaFree(plan->cycleTable);
        plan->cycleTable = nullptr;
    }
    if (cudaMalloc(&plan->cycleTable, residues.size() * sizeof(float)) != cudaSuccess || cudaMemcpy(plan->cycleTable, residues.data(), residues.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    plan->cycleMu = mu, plan->cycleLambda = lambda, plan->cycleStride = stride;
    plan->cycleValid = true;
}

unsigned long long cycleIndex(const gr4b200_rotator_plan* plan, unsigned long long m) {
    const unsigned long long size = plan->cycleMu + plan->cycleLambda;
    return m < size ? m : plan->cycleMu + (m - plan->cycleMu) % plan->cycleLambda;
}

} // namespace

namespace gr4b200 {
// used by the fused DDC path (ddc in fir.cu would need the per-run phases): fills plan->runPhases for a call of n samples
int rotatorPrepareCheckpoints(gr4b200_rotator_plan* plan, cudaStream_t stream, size_t n, const float** runPhases) {
    const size_t runs = ceilDiv<size_t>(n, kRun);
    if (runs > plan->runCapacity) {
        if (plan->runPhases != nullptr) {
            GR4B200_CUDA_TRY(cudaStreamSynchronize(stream));
            GR4B200_CUDA_TRY(cudaFree(plan->runPhases));
            plan->runPhases = nullptr;
        }
        GR4B200_CUDA_TRY(cudaMalloc(&plan->runPhases, runs * sizeof(float)));
        plan->runCapacity = runs;
    }
    if (!plan->cycleTried) {
        buildPhaseCycle(plan); // host replay, once per plan / set_phase
    }
    if (plan->cycleValid) {
        const unsigned long long size    = plan->cycleMu + plan->cycleLambda;
        gatherCheckpointKernel<<<static_cast<unsigned>(ceilDiv<unsigned long long>(runs, 256ull * kGatherPerThread)), 256, 0, stream>>>(plan->cycleTable, plan->cycleStride, plan->cycleMu, plan->cycleLambda, size, plan->position, runs, plan->runPhases);
        const unsigned long long end = cycleIndex(plan, plan->position + n);
        GR4B200_CUDA_TRY(cudaMemcpyAsync(plan->endPhase, plan->cycleTable + (end & 7ull) * plan->cycleStride + (end >> 3), sizeof(float), cudaMemcpyDeviceToDevice, stream));
        plan->pendingSamples = n;
        *runPhases           = plan->runPhases;
        return checkLaunch("rotator checkpoints (phase cycle)", 1u);
    }
    plan->pendingSamples = n;
    const int status     = ensureTables(plan, n, stream);
    if (status != GR4B200_OK) {
        return status;
    }
    if (plan->useTables) {
        prefixKernel<<<1, 1, 0, stream>>>(plan->landing, plan->phase, n, plan->prefix);
        // Samples per checkpoint thread. A thread pays ~log2(n) dependent table look-ups, replays up to ONE REVOLUTION
        // (2 pi / |dphi| steps) to reach its stretch and then the stretch itself, all serially. 512 is best for ordinary
        // increments (measured 128 / 256 / 512 at dphi = 0.63: 233.6 / 236.9 / 237.9 GS/s for the fused DDC,
        // profiles/r02x_time_mixer_checkpoints.jsonl); for small increments the revolution dominates (dphi = 1e-3: 6283 steps
        // in front of every 512-sample stretch, 140 GS/s), so the stretch grows with it -- fewer threads, each replaying one
        // revolution for a proportionally longer stretch. GR4B200_CHECKPOINT_STRETCH overrides.
        static const int forced = [] { const char* e = std::getenv("GR4B200_CHECKPOINT_STRETCH"); return e != nullptr ? std::atoi(e) : 0; }();
        const double     revolution = 6.283185307179586 / std::fabs(static_cast<double>(plan->dphi));
        const int        stretch    = forced > 0 ? forced : (revolution <= 1024.0 ? 512 : (revolution <= 4096.0 ? 2048 : 8192));
        const unsigned long long nStretches = ceilDiv<unsigned long long>(n, static_cast<unsigned long long>(stretch));
        const int                blocks     = static_cast<int>(ceilDiv<unsigned long long>(nStretches + 1, 128));
        if (stretch == 8192) {
            checkpointKernel<8192><<<blocks, 128, 0, stream>>>(plan->landing, plan->phase, plan->prefix, plan->tables, plan->nLevels, n, plan->runPhases, plan->endPhase);
        } else if (stretch == 2048) {
            checkpointKernel<2048><<<blocks, 128, 0, stream>>>(plan->landing, plan->phase, plan->prefix, plan->tables, plan->nLevels, n, plan->runPhases, plan->endPhase);
        } else if (stretch == 128) {
            checkpointKernel<128><<<blocks, 128, 0, stream>>>(plan->landing, plan->phase, plan->prefix, plan->tables, plan->nLevels, n, plan->runPhases, plan->endPhase);
        } else {
            checkpointKernel<512><<<blocks, 128, 0, stream>>>(plan->landing, plan->phase, plan->prefix, plan->tables, plan->nLevels, n, plan->runPhases, plan->endPhase);
        }
    } else {
        serialCheckpointKernel<<<1, 1, 0, stream>>>(plan->dphi, plan->phase, n, plan->runPhases, plan->endPhase, plan->settled);
        fillSettledKernel<<<static_cast<int>(std::min<unsigned long long>(ceilDiv<unsigned long long>(runs, 256), 4096)), 256, 0, stream>>>(plan->settled, n, plan->runPhases);
    }
    *runPhases = plan->runPhases;
    return checkLaunch("rotator checkpoints", 2u);
}
int rotatorCommitPhase(gr4b200_rotator_plan* plan, cudaStream_t stream) {
    plan->position += plan->pendingSamples; // the samples of the call whose checkpoints were prepared last
    plan->pendingSamples = 0;
    return checkCuda(cudaMemcpyAsync(plan->phase, plan->endPhase, sizeof(float), cudaMemcpyDeviceToDevice, stream), "rotator commit");
}
float rotatorIncrement(const gr4b200_rotator_plan* plan) { return plan->dphi; }
} // namespace gr4b200

extern "C" {

gr4b200_rotator_plan* gr4b200_rotator_plan_create(float phaseIncrement, float initialPhase) {
    auto* plan   = new gr4b200_rotator_plan;
    plan->device = currentDevice();
    plan->dphi   = phaseIncrement;
    bool ok    = cudaMalloc(&plan->phase, sizeof(float)) == cudaSuccess && cudaMalloc(&plan->endPhase, sizeof(float)) == cudaSuccess && cudaMalloc(&plan->prefix, sizeof(Prefix)) == cudaSuccess && cudaMalloc(&plan->failed, sizeof(int)) == cudaSuccess && cudaMalloc(&plan->settled, sizeof(Settled)) == cudaSuccess;
    ok         = ok && cudaMemcpy(plan->phase, &initialPhase, sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        checkCuda(cudaGetLastError(), "rotator_plan_create");
        gr4b200_rotator_plan_destroy(plan);
        return nullptr;
    }
    plan->useTables  = landingFor(phaseIncrement, plan->landing);
    plan->startPhase = initialPhase;
    return plan;
}

int gr4b200_rotator_plan_destroy(gr4b200_rotator_plan* plan) {
    if (plan == nullptr) {
        return GR4B200_OK;
    }
    cudaFree(plan->phase);
    cudaFree(plan->endPhase);
    cudaFree(plan->prefix);
    cudaFree(plan->failed);
    cudaFree(plan->settled);
    cudaFree(plan->tables);
    cudaFree(plan->runPhases);
    cudaFree(plan->cycleTable);
    delete plan;
    return GR4B200_OK;
}

int gr4b200_rotator_set_phase(gr4b200_rotator_plan* plan, float accumulatedPhase) {
    if (plan == nullptr) {
        return fail("rotator_set_phase: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "rotator_set_phase"); status != GR4B200_OK) {
        return status;
    }
    GR4B200_CUDA_TRY(cudaDeviceSynchronize()); // the plan's device: launches that still read the old phase finish first
    plan->startPhase = accumulatedPhase;       // a new origin for the phase cycle, rebuilt on the next call
    plan->position   = 0;
    plan->cycleTried = false;
    plan->cycleValid = false;
    return checkCuda(cudaMemcpy(plan->phase, &accumulatedPhase, sizeof(float), cudaMemcpyHostToDevice), "rotator_set_phase");
}

float gr4b200_rotator_get_phase(const gr4b200_rotator_plan* plan) {
    float phase = NAN;
    if (plan != nullptr && checkPlanDevice(plan->device, "rotator_get_phase") == GR4B200_OK) {
        cudaDeviceSynchronize();
        cudaMemcpy(&phase, plan->phase, sizeof(float), cudaMemcpyDeviceToHost);
    }
    return phase;
}

float gr4b200_rotator_phase_increment(float frequencyShift, float sampleRate) { return 2.f * static_cast<float>(3.14159265358979323846f * frequencyShift / sampleRate); }

int gr4b200_rotator_cf32(gr4b200_rotator_plan* plan, void* stream, const float* in, float* out, size_t n) {
    if (plan == nullptr) {
        return fail("rotator: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "rotator"); status != GR4B200_OK) {
        return status;
    }
    if (n == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("rotator: null or misaligned buffer");
    }
    const auto   s         = asStream(stream);
    const float* runPhases = nullptr;
    int          status    = rotatorPrepareCheckpoints(plan, s, n, &runPhases);
    if (status != GR4B200_OK) {
        return status;
    }
    const unsigned long long nTiles = ceilDiv<unsigned long long>(n, kTile);
    // one CTA per tile (no resident grid): a streaming kernel is faster under the hardware CTA scheduler, see mathop.cu;
    // GR4B200_ROTATOR_CTAS=n restores a resident grid of n CTAs per SM for A/B timing
    static const int         residentCtas = [] { const char* e = std::getenv("GR4B200_ROTATOR_CTAS"); return e != nullptr ? std::atoi(e) : 0; }();
    const unsigned long long cap          = residentCtas > 0 ? static_cast<unsigned long long>(smCount()) * residentCtas : nTiles;
    const int  grid  = static_cast<int>(nTiles < cap ? nTiles : cap);
    const bool signed_ = std::isfinite(plan->dphi) && plan->dphi != 0.f && std::fabs(plan->dphi) <= 3.1415927f;
    if (signed_ && plan->dphi > 0.f) {
        rotateKernel<1><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), n, plan->dphi, runPhases);
    } else if (signed_) {
        rotateKernel<-1><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), n, plan->dphi, runPhases);
    } else {
        rotateKernel<0><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), n, plan->dphi, runPhases);
    }
    status = checkLaunch("rotateKernel");
    if (status != GR4B200_OK) {
        return status;
    }
    return rotatorCommitPhase(plan, s);
}

} // extern "C"
