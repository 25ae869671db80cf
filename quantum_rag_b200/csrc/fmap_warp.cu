// K1b: amplitude-encoded state followed by L feature-map layers at n = 10 qubits
// (BASELINE config 5: 1024-d embeddings), one WARP per state, complex128.
//
// What it replaces: the per-pair `_vector_to_circuit` + two statevector runs of
// /root/reference/src/reranker/quantum.py:108-167, with the amplitude-encoded
// start state quantum.py:156 names and the layer block repeated (builder-defined,
// SURVEY.md section 8d config 5; oracle: oracle/quantum.py feature_map_state).
//
// Why a second kernel beside sv_cta_kernel: that one keeps the state in shared
// memory and makes four radix-8 round trips per layer with a CTA barrier after
// each, applying RY and RZ as a general complex 2x2 per qubit (80 flops per
// amplitude and layer).  Here
//   * a lane holds 32 of the 1024 amplitudes in REGISTERS; five qubits are
//     register-index bits, so a layer is two passes (qubits 0-4 and 5-9) with ONE
//     warp-synchronous transpose through shared memory between them (even layers
//     run 0-4 first, odd layers 5-9 first) - no CTA barrier on the path;
//   * the layer is regrouped as [all RY][all RZ] (gates on different qubits
//     commute).  RY = c [[1, -t], [t, 1]] is a scaled real rotation: 2 FMAs per
//     amplitude, the product of the c's goes into the phase table (used when every
//     |t| <= 1, i.e. every |x^_i| <= 1/2; otherwise the direct 4-flop form).  The
//     ten RZ collapse into one diagonal per layer = table[register index] x
//     constant[lane], one table load and 8 flops per amplitude;
//   * layer 0 stays real until its diagonal;
//   * the CX chain between layers is a register renaming plus, per side, either a
//     conditional register reversal or one 16-amplitude __shfl_xor with lane ^ 1
//     (see fw_layers); the chain of the LAST layer is dropped on both sides of the
//     overlap (the same permutation of both states);
//   * the query state is evolved once per (query, chunk of candidates) by warp 0
//     while the other warps already evolve their first candidate.
// Shared-memory layout of a state: element (row, col) at row * 33 + col (one unit of padding per
// 32): row-wise and column-wise accesses are bank-conflict free and every access is a
// lane-dependent base plus a compile-time offset (no address registers).
//
// Bound: the FP64 pipe (64 FMA/clk/SM).  Per state and layer 10 x 2 + 8 = 28 FP64 instructions
// per amplitude (896 per lane); measured 3.7e3 FP64 warp instructions per state at L = 4 against
// 6.2e3 for the direct form; HBM traffic is the 4 KB row.  ncu: profiles/r01_fmap_warp_*.
#include "common.cuh"

namespace qrag {

namespace {

constexpr int FW_N = 10;
constexpr int FW_DIM = 1 << FW_N;
constexpr int FW_ST = FW_DIM + 32;                          // state stride: rows of 32 amplitudes + 1 of padding
constexpr int FW_STAGE = FW_DIM + (FW_DIM >> 5) * 4;        // fp32 row staging: 4 floats of padding per 32

struct FmapWarpParams {
    const float* Q; const float* cand; const float* X; const int64_t* idx;
    int64_t N;
    int64_t C; int D; int nq; int layers;
    int64_t chunk; int chunks_per_q; int64_t units;
    double* out; float* out32;
};

// The state's scalar type: double (exact path, complex128) or float (the FILTER pass of the rerank: complex64 evolution,
// fp64 overlap; its error bound and the certification of the top-k boundary are in fmap_rerank.cu).
template <typename R> struct Real2;
template <> struct Real2<double> { using type = double2; };
template <> struct Real2<float> { using type = float2; };
template <typename R> using R2 = typename Real2<R>::type;
template <typename R> __device__ __forceinline__ R2<R> make_r2(R x, R y) { R2<R> v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ double shfl_xor_r(double v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }
__device__ __forceinline__ float shfl_xor_r(float v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }

template <typename R> struct FwGate { R c, s, cp, sp; };      // RY half-angle cos/sin, RZ half-angle cos/sin

__host__ __device__ constexpr int fw_pxor5(int x) {          // CX chain on five bits: y_k = x_0 ^ ... ^ x_k
    x ^= x << 1; x ^= x << 2; x ^= x << 4;
    return x & 31;
}
__host__ __device__ constexpr int fw_parity5(int x) { return (x ^ (x >> 1) ^ (x >> 2) ^ (x >> 3) ^ (x >> 4)) & 1; }

// RY on the five qubits that are bits of the register index.
// FAST: RY = c * [[1, -t], [t, 1]], t = tan(half angle); the factor c goes into the phase table, so a
// complex pair costs 4 FMAs instead of 8 flops.  Used when every |t| <= 1 (|a| <= 1/2), else the direct form.
template <bool FAST, typename R>
__device__ __forceinline__ void fw_ry5(R2<R> (&a)[32], const FwGate<R>* __restrict__ g, const R* __restrict__ tn) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        if (FAST) {
            const R t = tn[b];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j & (1 << b)) continue;
                const int k = j | (1 << b);
                const R2<R> a0 = a[j], a1 = a[k];
                a[j].x = fma(-t, a1.x, a0.x);
                a[j].y = fma(-t, a1.y, a0.y);
                a[k].x = fma(t, a0.x, a1.x);
                a[k].y = fma(t, a0.y, a1.y);
            }
        } else {
            const R c = g[b].c, s = g[b].s;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j & (1 << b)) continue;
                const int k = j | (1 << b);
                const R2<R> a0 = a[j], a1 = a[k];
                a[j].x = fma(-s, a1.x, c * a0.x);
                a[j].y = fma(-s, a1.y, c * a0.y);
                a[k].x = fma(s, a0.x, c * a1.x);
                a[k].y = fma(s, a0.y, c * a1.y);
            }
        }
    }
}

template <bool FAST, typename R>
__device__ __forceinline__ void fw_ry5_real(R (&r)[32], const FwGate<R>* __restrict__ g,
                                            const R* __restrict__ tn) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const R t = tn[b], c = g[b].c, s = g[b].s;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j & (1 << b)) continue;
            const int k = j | (1 << b);
            const R a0 = r[j], a1 = r[k];
            if (FAST) {
                r[j] = fma(-t, a1, a0);
                r[k] = fma(t, a0, a1);
            } else {
                r[j] = fma(-s, a1, c * a0);
                r[k] = fma(s, a0, c * a1);
            }
        }
    }
}

// Phase table of five RZ gates: entry j = prod_b (bit b of j ? e^{+i phi_b/2} : e^{-i phi_b/2}),
// times the RY scale prod_b c_b in the FAST form.
template <bool FAST, typename R>
__device__ __forceinline__ R2<R> fw_phase(const FwGate<R>* __restrict__ g, int j) {
    // evaluated in R: for R = float that is 5 complex products of rounded unit factors (<= 4u each) and, in the
    // scaled form, 5 real factors -- accounted for in the filter's error bound (fmap_rerank.cu)
    R2<R> p = make_r2<R>((R)1, (R)0);
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const R cr = g[b].cp;
        const R ci = ((j >> b) & 1) ? g[b].sp : -g[b].sp;
        const R nx = p.x * cr - p.y * ci;
        const R ny = p.x * ci + p.y * cr;
        p.x = nx; p.y = ny;
    }
    if (FAST) {
        const R sc = g[0].c * g[1].c * g[2].c * g[3].c * g[4].c;
        p.x *= sc; p.y *= sc;
    }
    return p;
}

// a[j] *= tab[j] * cst: the RZ diagonal of a whole layer, split as (uniform table over the register
// index) x (per-lane constant for the lane's own row / column).
template <typename R>
__device__ __forceinline__ void fw_diag(R2<R> (&a)[32], const R2<R>* __restrict__ tab, const R2<R> cst) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const R2<R> p = tab[j], v = a[j];
        const R wx = p.x * cst.x - p.y * cst.y;
        const R wy = fma(p.x, cst.y, p.y * cst.x);
        a[j].x = v.x * wx - v.y * wy;
        a[j].y = fma(v.x, wy, v.y * wx);
    }
}

// The layers.  Qubits 0-4 are the low five bits of the basis index x, qubits 5-9 the high five.
//   layout A: a lane holds one value of the high bits (its "row") and all 32 low values in registers;
//   layout B: a lane holds one value of the low bits (its "column") and all 32 high values.
// A pass rotates the five qubits that are register bits.  Even layers run A then B, odd layers B
// then A, so there is ONE transpose through shared memory per layer, and the layer's RZ diagonal
// is applied once, after the second pass (it commutes with the RY of the other five qubits):
// table[register index] x constant[lane's own row or column].  Layer 0 stays REAL until that
// diagonal.  The CX chain between layers (new[y] = old[x], y = prefix-xor of x) never touches
// shared memory:
//   after a B pass (column l, registers = high bits h): y_lo = pxor5(l) is a new column label,
//   y_hi = pxor5(h) ^ (parity(l) ? 31 : 0) a register renaming plus, on odd-parity lanes, a reversal;
//   after an A pass (row h, registers = low bits l): y_lo = pxor5(l) is a renaming, and y_hi =
//   pxor5(h) ^ (parity(l) ? 31 : 0) means the odd-parity registers belong to the row of lane ^ 1:
//   one __shfl_xor of 16 amplitudes.
// Ends in layout B (layers odd) or A (layers even); the same for both states of an overlap.
template <bool FAST, typename R>
__device__ __forceinline__ void fw_layers(R (&r)[32], R2<R> (&a)[32], int layers, R2<R>* __restrict__ st,
                                          const FwGate<R>* __restrict__ gates, const R* __restrict__ tn,
                                          R2<R>* __restrict__ tab) {
    const int lane = threadIdx.x & 31;
    const bool odd_lane = __popc(lane) & 1;
    const int lcol = fw_pxor5(lane);                         // column label after a B-side CX
    int hrow = lane;                                         // row label (changes after an A-side CX)
    // ---- layer 0 on the REAL amplitudes: A, transpose, B, then the diagonal makes them complex
    {
        tab[lane] = fw_phase<FAST, R>(gates + 5, lane);
        const R2<R> cst = fw_phase<FAST, R>(gates, lane);
        fw_ry5_real<FAST, R>(r, gates, tn);
        R* sr = reinterpret_cast<R*>(st);
#pragma unroll
        for (int j = 0; j < 32; ++j) sr[lane * 33 + j] = r[j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = sr[j * 33 + lane];
        fw_ry5_real<FAST, R>(r, gates + 5, tn + 5);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const R2<R> p = tab[j];
            a[j].x = r[j] * (p.x * cst.x - p.y * cst.y);
            a[j].y = r[j] * fma(p.x, cst.y, p.y * cst.x);
        }
        __syncwarp();
    }
    // one generic layer body for layers >= 1 (about 1 300 instructions; two parity-specialised copies measured the
    // same speed): which five qubits come first, the transpose direction and the CX side follow the layer's parity at run time
    for (int layer = 1; layer < layers; ++layer) {
        const bool odd = layer & 1;
        if (odd) {   // B-side CX chain, in registers
            R2<R> b[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) b[fw_pxor5(j)] = a[j];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                a[j].x = odd_lane ? b[31 - j].x : b[j].x;
                a[j].y = odd_lane ? b[31 - j].y : b[j].y;
                a[31 - j].x = odd_lane ? b[j].x : b[31 - j].x;
                a[31 - j].y = odd_lane ? b[j].y : b[31 - j].y;
            }
        } else {     // A-side CX chain: odd-parity registers come from lane ^ 1, then the renaming
            R2<R> b[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                R2<R> v = a[j];
                if (fw_parity5(j)) {
                    v.x = shfl_xor_r(v.x, 1);
                    v.y = shfl_xor_r(v.y, 1);
                }
                b[fw_pxor5(j)] = v;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = b[j];
            hrow = lcol;
        }
        const FwGate<R>* g = gates + layer * FW_N;
        const R* t = tn + layer * FW_N;
        const int o1 = odd ? 5 : 0, o2 = 5 - o1;
        tab[lane] = fw_phase<FAST, R>(g + o2, lane);
        const R2<R> cst = fw_phase<FAST, R>(g + o1, lane);
        fw_ry5<FAST, R>(a, g + o1, t + o1);
        R2<R>* sp = st + (odd ? lcol : hrow * 33);
        const int ss = odd ? 33 : 1;
#pragma unroll
        for (int j = 0; j < 32; ++j) sp[j * ss] = a[j];
        __syncwarp();
        const R2<R>* lp = st + (odd ? lane * 33 : lane);
        const int ls = odd ? 1 : 33;
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = lp[j * ls];
        fw_ry5<FAST, R>(a, g + o2, t + o2);
        fw_diag<R>(a, tab, cst);
        __syncwarp();                                    // reads of st and tab are complete
    }
}

// Evolves one fp32 row into the state after `layers` blocks WITHOUT the last CX chain, in the
// final register layout of fw_layers.  Returns |row| == 0.  smem regions are private to the warp.
__device__ __forceinline__ bool fw_row_vec(const float* row, int D) {
    return ((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
}

// Asynchronous copy of a (16-byte aligned, D % 4 == 0) row into the warp's padded staging buffer, zero-filled past
// D: issued while the previous state is still being evolved, so a state does not start with an HBM round trip.
__device__ __forceinline__ void fw_prefetch_row(const float* __restrict__ row, int D, float* __restrict__ stage) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < FW_DIM / 128; ++r) {
        const int m = lane + 32 * r;                         // float4 index
        const bool in = 4 * m < D;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + 4 * m + (m >> 3) * 4);
        const float* src = in ? row + 4 * m : row;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// `stage` holds the row if `staged` (prefetched by the previous call), else it is loaded here.  `next_row` (or
// nullptr) is prefetched into `stage` as soon as this row has been consumed; returns through `*next_staged`.
template <typename R>
__device__ __forceinline__ bool fw_evolve(const float* __restrict__ row, int D, int layers, R2<R> (&a)[32],
                                       R2<R>* __restrict__ st, float* __restrict__ stage, bool staged,
                                       const float* __restrict__ next_row, bool* next_staged,
                                       FwGate<R>* __restrict__ gates, R* __restrict__ tn, R2<R>* __restrict__ tab) {
    const int lane = threadIdx.x & 31;
    // ---- row -> padded staging (coalesced; 4 floats of padding per 32)
    const bool vec = fw_row_vec(row, D);
    if (staged) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (vec) {
#pragma unroll
        for (int r = 0; r < FW_DIM / 128; ++r) {
            const int m = lane + 32 * r;                     // float4 index
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * m < D) v = ldg_stream(reinterpret_cast<const float4*>(row) + m);
            *reinterpret_cast<float4*>(stage + 4 * m + (m >> 3) * 4) = v;
        }
    } else {
        for (int i = lane; i < FW_DIM; i += 32) stage[i + (i >> 5) * 4] = i < D ? row[i] : 0.f;
    }
    __syncwarp();
    // ---- my 32 real amplitudes (basis index (lane << 5) | j), |row|^2 in fp64
    double rd[32];
    double n2 = 0.0;
    {
        const float4* src = reinterpret_cast<const float4*>(stage + lane * 36);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 v = src[u];
            rd[4 * u + 0] = (double)v.x; rd[4 * u + 1] = (double)v.y;
            rd[4 * u + 2] = (double)v.z; rd[4 * u + 3] = (double)v.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) n2 = fma(rd[j], rd[j], n2);
    }
    n2 = warp_sum(n2);
    const double nrm = sqrt(n2);
    const bool zero = !(nrm > 0.0);
    const double inv = zero ? 1.0 : 1.0 / nrm;              // one division per state; x * inv is within 1 ulp of x / nrm
    R r[32];                                                 // (normalised in double, then rounded to R once)
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = (R)(rd[j] * inv);
    // ---- gate parameters: layer l, qubit i takes a = x^[(l*n + i) % D]  (quantum.py:160-161: ry(a pi), rz(a pi/2));
    // evaluated in double whatever R is (a few dozen per state), rounded to R once
    bool small = true;
    for (int t = lane; t < layers * FW_N; t += 32) {
        const int comp = t % D;
        const double an = (double)stage[comp + (comp >> 5) * 4] * inv;
        double sp, cp;
        sincospi(0.25 * an, &sp, &cp);                       // RZ half angle a pi / 4; the RY half angle is twice that:
        const double sn = 2.0 * sp * cp;                     // sin 2x = 2 sin x cos x
        const double cs = fma(-2.0 * sp, sp, 1.0);           // cos 2x = 1 - 2 sin^2 x (|x| <= pi/4: no cancellation
        FwGate<R> g;                                         //  beyond 1e-16 absolute)
        g.sp = (R)sp; g.cp = (R)cp; g.s = (R)sn; g.c = (R)cs;
        gates[t] = g;
        tn[t] = (R)(sn / cs);
        small = small && (fabs(an) <= 0.5);
    }
    const bool fast = __all_sync(FULL_MASK, small);
    __syncwarp();                                            // gates visible; staging fully consumed
    *next_staged = next_row != nullptr && fw_row_vec(next_row, D);
    if (*next_staged) fw_prefetch_row(next_row, D, stage);
    if (fast) fw_layers<true, R>(r, a, layers, st, gates, tn, tab);
    else      fw_layers<false, R>(r, a, layers, st, gates, tn, tab);
    return zero;
}

// One CTA = W warps; a unit = (query, chunk of its candidates).  Item 0 of a unit is the query
// itself, item t >= 1 candidate t-1; warp w takes items w, w + W, ...  Warp 0 evolves the query
// state while the others already evolve their first candidate.
template <int MAXT, typename R>
__global__ void __launch_bounds__(MAXT, 1) fmap_warp_kernel(const FmapWarpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout (the same order as fmap_warp_smem): query state | W states | W tables | W x layers x 10 gates | tangents | staging
    R2<R>* qstate = reinterpret_cast<R2<R>*>(smem_raw);
    R2<R>* st = qstate + FW_ST + (size_t)warp * FW_ST;
    R2<R>* tab = qstate + FW_ST + (size_t)W * FW_ST + warp * 32;
    FwGate<R>* gates0 = reinterpret_cast<FwGate<R>*>(qstate + FW_ST + (size_t)W * (FW_ST + 32));
    FwGate<R>* gates = gates0 + warp * p.layers * FW_N;
    R* tn0 = reinterpret_cast<R*>(gates0 + (size_t)W * p.layers * FW_N);
    R* tn = tn0 + warp * p.layers * FW_N;
    // (W * layers * 10 tangents; the staging rows need 16-byte alignment: round the float count up to 4)
    float* stage = reinterpret_cast<float*>(tn0 + (((size_t)W * p.layers * FW_N + 3) & ~(size_t)3)) + (size_t)warp * FW_STAGE;
    __shared__ int q_zero;

    R2<R> a[32];
    for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int64_t qi = u / p.chunks_per_q;
        const int64_t c0 = (u % p.chunks_per_q) * p.chunk;
        int64_t cnt = p.C - c0;
        if (cnt > p.chunk) cnt = p.chunk;
        // every warp runs every round (uniform trip count).  Starting each round with a barrier, so that the warps
        // stream the same straight-line code together and share instruction fetches, was measured: no gain.
        bool staged = false;                                 // the first row of a unit is loaded directly
        auto row_of = [&](int64_t t, bool* missing) -> const float* {
            *missing = false;
            if (t == 0) return p.Q + qi * p.D;
            const int64_t j = qi * p.C + c0 + (t - 1);
            if (p.cand) return p.cand + (size_t)j * p.D;
            const int64_t id = p.idx[j];
            *missing = id < 0 || id >= p.N;
            return p.X + (size_t)(*missing ? 0 : id) * p.D;
        };
        for (int64_t t = warp; t - warp <= cnt; t += W) {
            const bool active = t <= cnt;
            const int64_t j = qi * p.C + c0 + (t - 1);
            bool missing = false, zero = false;
            if (active) row_of(t, &missing);
            if (active && missing) {                         // padding candidate (warp-uniform): no state to evolve
                if (staged) asm volatile("cp.async.wait_group 0;" ::: "memory");
                staged = false;
            } else if (active) {
                const float* row = row_of(t, &missing);
                bool nmiss = false;
                const float* next_row = (t + W <= cnt) ? row_of(t + W, &nmiss) : nullptr;
                bool next_staged = false;
                zero = fw_evolve<R>(row, p.D, p.layers, a, st, stage, staged, next_row, &next_staged, gates, tn, tab);
                staged = next_staged;
                if (t == 0) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) qstate[k * 33 + lane] = a[k];
                    if (lane == 0) q_zero = zero;
                }
            }
            if (t == warp) __syncthreads();                  // round 0: the query state of this unit is in place
            if (active && t > 0) {
                double re = 0.0, im = 0.0;                   // <psi_d | psi_q> = sum conj(d) q
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const R2<R> qr = qstate[k * 33 + lane];          // fp64 accumulation of the overlap whatever R is
                    const double qx = (double)qr.x, qy = (double)qr.y, dx = (double)a[k].x, dy = (double)a[k].y;
                    re = fma(dx, qx, re); re = fma(dy, qy, re);
                    im = fma(dx, qy, im); im = fma(-dy, qx, im);
                }
                re = warp_sum(re);
                im = warp_sum(im);
                if (lane == 0) {
                    double f = re * re + im * im;
                    if (q_zero || zero) f = 0.0;
                    if (missing) f = -__longlong_as_double(0x7ff0000000000000LL);
                    p.out[j] = f;
                    if (p.out32) p.out32[j] = (float)f;
                }
            }
        }
        __syncthreads();                                     // everyone is done with qstate before the next unit
    }
}


template <typename R>
static size_t fmap_warp_smem(int W, int layers) {
    const size_t tn = (((size_t)W * layers * FW_N + 3) & ~(size_t)3);
    return (size_t)(1 + W) * FW_ST * sizeof(R2<R>) + (size_t)W * 32 * sizeof(R2<R>) +
           (size_t)W * layers * FW_N * sizeof(FwGate<R>) + tn * sizeof(R) + (size_t)W * FW_STAGE * sizeof(float);
}

template <typename R, int MAXW>
static int fmap_warp_launch(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C,
                            int D, int layers, double* out64, float* out32, cudaStream_t st, bool* handled) {
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    // complex128: 8 warps = 2 per scheduler at 255 registers.  Measured on B200 (512 x 1000 x 1024, L = 4): 7.9e7 scores/s;
    // 11 warps at 168 registers 6.6e7 (spills, unbalanced schedulers); the 168-register build at 8 warps 6.2e7.
    // complex64 (the rerank's filter pass): half the registers and half the shared memory per state, 12 warps.
    int W = MAXW;
    while (W > 1 && fmap_warp_smem<R>(W, layers) + 64 > (size_t)dp.max_smem_optin) --W;
    if (fmap_warp_smem<R>(W, layers) + 64 > (size_t)dp.max_smem_optin) return QRAG_OK;    // layers too deep: generic kernel
    FmapWarpParams p{};
    p.Q = Q; p.cand = cand; p.X = X; p.idx = idx; p.N = N; p.C = C; p.D = D; p.nq = nq; p.layers = layers;
    p.out = out64; p.out32 = out32;
    // a unit should fill at least one round of the CTA (query + W-1 candidates); prefer >= 2 units per SM
    int64_t cpq = ceil_div(2 * (int64_t)dp.sm_count, nq);
    const int64_t max_cpq = ceil_div(C, W > 1 ? W - 1 : 1);
    if (cpq > max_cpq) cpq = max_cpq;
    if (cpq < 1) cpq = 1;
    p.chunk = ceil_div(C, cpq);
    p.chunks_per_q = (int)ceil_div(C, p.chunk);
    p.units = (int64_t)nq * p.chunks_per_q;
    const int64_t grid = p.units < dp.sm_count ? p.units : dp.sm_count;
    const size_t smem = fmap_warp_smem<R>(W, layers);
    auto kern = fmap_warp_kernel<MAXW * 32, R>;
    QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, W * 32, smem, st>>>(p);
    QRAG_LAUNCH_CHECK("fmap_warp_kernel");
    *handled = true;
    return QRAG_OK;
}

}  // namespace

// Returns QRAG_OK and sets *handled when the shape is served by this kernel (n = 10).
int fmap_warp_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                  int n_qubits, int layers, double* out64, float* out32, cudaStream_t st, bool* handled) {
    *handled = false;
    if (n_qubits != FW_N || layers < 1 || D < 1 || D > FW_DIM) return QRAG_OK;
    return fmap_warp_launch<double, 8>(Q, nq, cand, X, N, idx, C, D, layers, out64, out32, st, handled);
}

// The same evolution with the state in complex64 (fp64 normalisation, gate parameters and overlap): the FILTER pass of
// qrag_fmap_rerank.  |F32 - F| <= fmap_filter_error_bound(layers) (fmap_rerank.cu); never a result by itself.
int fmap_warp_filter_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C,
                         int D, int n_qubits, int layers, double* out64, cudaStream_t st, bool* handled) {
    *handled = false;
    if (n_qubits != FW_N || layers < 1 || D < 1 || D > FW_DIM) return QRAG_OK;
    return fmap_warp_launch<float, 12>(Q, nq, cand, X, N, idx, C, D, layers, out64, nullptr, st, handled);
}

}  // namespace qrag
