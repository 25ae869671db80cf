// K1, register-resident form: exact complex128 statevector of the reference fidelity circuit
// (QuantumReranker._quantum_similarity / _vector_to_circuit, /root/reference/src/reranker/quantum.py:108-167)
// for n_qubits <= 10.  n = 11, 12 stay on the shared-memory kernel in sv_kernels.cu.
//
// What the reference does per (query, document) pair: build two circuits, run the simulator twice, take
// |<psi_d|psi_q>|^2.  Here, per batch:
//   pass 1  every QUERY state once            -> scratch [nq][2^n] complex128 (stream-ordered pool allocation)
//   pass 2  every DOCUMENT state, in registers -> overlap with its query's state -> out[j]
// Pass 2 is launched with programmatic stream serialisation and builds its first state before it waits for
// pass 1, so the two overlap.
//
//  * n <= 5: one THREAD per state, the 2^n amplitudes in registers (all indices static: gates are register
//    arithmetic, the CX chain a register renaming).  No shuffles, no redundant trigonometry: every thread owns a
//    different vector.
//  * 6 <= n <= 10: one WARP per state, 2^(n-5) amplitudes per lane (qubits 0..n-6 inside a lane, the top five
//    across lanes).  Lane k evaluates the gate of qubit k (one sincospi pair per gate per state; the first block's
//    gates of TWO states per pass, one state per half-warp) and the warp shares it by shuffles; the CX chain is one shuffle per amplitude (source register static, source lane
//    lane ^ (lane << 1) ^ carry).
//
// The first block of gates acts on |0...0>: when qubit k's RY/RZ is applied, every amplitude with a bit >= k set is
// still exactly zero, so the gate maps old[x] -> (alpha_k old[x], beta_k old[x]) for x < 2^k and touches nothing
// else.  Those zero amplitudes are skipped (an exact identity, not an approximation): layer 0 costs 2 (2^n - 1)
// complex multiplications instead of n 2^(n-1) dense 2x2 updates.  Layers >= 1 (`layers` > 1, the builder's
// extension) are dense.  The full 2^n-amplitude state is always materialised and the CX chain applied to it.
//
// Gate semantics (Qiskit, little-endian): RY(t) = [[c,-s],[s,c]], c = cos t/2; RZ(p) = diag(e^{-ip/2}, e^{+ip/2});
// quantum.py:160-161: t = a pi, p = a pi / 2, so the half angles are a pi/2 and a pi/4 -- evaluated by sincospi,
// which needs no range reduction against a rounded pi.  CX(i,i+1), i = 0..n-2: new[y] = old[y ^ (y << 1)].
#include "common.cuh"
#include "tma.cuh"

#include <mutex>

namespace qrag {

struct SvaParams {
    const double* qvec; const double* dvec; const int32_t* doc_query;
    int64_t nd, docs_per_query; int nq, vec_len, layers;
    double2* qstate;      // scratch [nq][2^n]; thread form: natural order; warp form: [register][lane]
    double* out;
};

constexpr int SVA_WARP_THREADS = 128;      // warp form: 4 states per CTA (registers, not threads, bound the occupancy)

struct SvaGate { double c, s, cp, sp; };   // RY half-angle cos/sin, RZ half-angle cos/sin

__device__ __forceinline__ SvaGate sva_gate(double a) {
    SvaGate g;
    sincospi(0.5 * a, &g.s, &g.c);
    sincospi(0.25 * a, &g.sp, &g.cp);
    return g;
}

// normalised component that drives qubit k in `layer` (0 = not rotated: quantum.py:158 stops at min(len, n))
__device__ __forceinline__ double sva_angle(const double* __restrict__ v, int vec_len, double nrm, int n, int layers,
                                            int layer, int k) {
    const int limit = vec_len < n ? vec_len : n;
    if (layers == 1 && k >= limit) return 0.0;
    const double x = v[(layer * n + k) % vec_len];
    return nrm > 0.0 ? x / nrm : x;                        // quantum.py:149-151
}

__device__ __forceinline__ double sva_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// RZ(phi) RY(theta) on the amplitude pair (a0 = bit clear, a1 = bit set)
__device__ __forceinline__ void sva_apply(const SvaGate& g, double2& a0, double2& a1) {
    const double t0r = g.c * a0.x - g.s * a1.x, t0i = g.c * a0.y - g.s * a1.y;
    const double t1r = g.s * a0.x + g.c * a1.x, t1i = g.s * a0.y + g.c * a1.y;
    a0.x = t0r * g.cp + t0i * g.sp;  a0.y = t0i * g.cp - t0r * g.sp;
    a1.x = t1r * g.cp - t1i * g.sp;  a1.y = t1i * g.cp + t1r * g.sp;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }
__device__ __forceinline__ double2 shfl_d2(double2 v, int src) { return make_double2(shfl_d(v.x, src), shfl_d(v.y, src)); }

// ===========================================================================
// thread form, n <= 5
// ===========================================================================
template <int N>
__device__ __forceinline__ void thread_cx(double2 (&st)[1 << N]) {
    constexpr int DIM = 1 << N;
    double2 t[DIM];
#pragma unroll
    for (int y = 0; y < DIM; ++y) t[y] = st[(y ^ (y << 1)) & (DIM - 1)];
#pragma unroll
    for (int y = 0; y < DIM; ++y) st[y] = t[y];
}

template <int N, bool MULTI>
__device__ __forceinline__ void thread_state(const double* __restrict__ v, int vec_len, int layers,
                                             double2 (&st)[1 << N]) {
    constexpr int DIM = 1 << N;
    double n2 = 0.0;
    for (int i = 0; i < vec_len; ++i) n2 = fma(v[i], v[i], n2);
    const double nrm = sqrt(n2);
    st[0] = make_double2(1.0, 0.0);
#pragma unroll
    for (int k = 0; k < N; ++k) {                          // layer 0 on |0..0>: only x < 2^k is populated
        const SvaGate g = sva_gate(sva_angle(v, vec_len, nrm, N, layers, 0, k));
        const double2 al = make_double2(g.c * g.cp, -g.c * g.sp), be = make_double2(g.s * g.cp, g.s * g.sp);
#pragma unroll
        for (int x = 0; x < (1 << k); ++x) {
            st[x | (1 << k)] = cmul(be, st[x]);
            st[x] = cmul(al, st[x]);
        }
    }
    thread_cx<N>(st);
    if (!MULTI) return;                                    // the reference's circuit is one block (quantum.py:158-165)
    for (int layer = 1; layer < layers; ++layer) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const SvaGate g = sva_gate(sva_angle(v, vec_len, nrm, N, layers, layer, k));
#pragma unroll
            for (int x = 0; x < DIM; ++x)
                if (!((x >> k) & 1)) sva_apply(g, st[x], st[x | (1 << k)]);
        }
        thread_cx<N>(st);
    }
}

template <int N, bool QUERY, bool MULTI>
__global__ void __launch_bounds__(128) sva_thread_kernel(const SvaParams p) {
    constexpr int DIM = 1 << N;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (QUERY) griddep_launch_dependents();
    if (j >= (QUERY ? (int64_t)p.nq : p.nd)) return;
    double2 st[DIM];
    thread_state<N, MULTI>((QUERY ? p.qvec : p.dvec) + j * p.vec_len, p.vec_len, p.layers, st);
    if (QUERY) {
#pragma unroll
        for (int y = 0; y < DIM; ++y) p.qstate[j * DIM + y] = st[y];
        return;
    }
    const int64_t qi = p.doc_query ? (int64_t)p.doc_query[j] : j / p.docs_per_query;
    if (qi < 0 || qi >= p.nq) { p.out[j] = sva_nan(); return; }   // a doc_query entry that names no query: never an OOB read
    griddep_wait();                                        // the query states are pass 1's output
    const double2* __restrict__ q = p.qstate + qi * DIM;
    double re = 0.0, im = 0.0;                             // <psi_d|psi_q> = sum conj(d) q
#pragma unroll
    for (int y = 0; y < DIM; ++y) {
        const double2 d = st[y], a = q[y];
        re = fma(d.x, a.x, fma(d.y, a.y, re));
        im = fma(d.x, a.y, fma(-d.y, a.x, im));
    }
    p.out[j] = re * re + im * im;
}

// ===========================================================================
// warp form, 6 <= n <= 10
// ===========================================================================
template <int N>
__device__ __forceinline__ void warp_cx(double2 (&st)[1 << (N - 5)], int lane) {
    constexpr int M = N - 5, R = 1 << M;
    const int src0 = (lane ^ (lane << 1)) & 31;
    double2 t[R];
#pragma unroll
    for (int ry = 0; ry < R; ++ry) {
        const int rs = (ry ^ (ry << 1)) & (R - 1);         // source register: static
        const int carry = (ry >> (M - 1)) & 1;             // qubit M-1 feeds lane bit 0
        t[ry] = shfl_d2(st[rs], src0 ^ carry);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) st[r] = t[r];
}

// Gates of the first block for TWO states at once: lanes 0..15 serve state A, lanes 16..31 state B (n <= 10 < 16 gates
// per state).  The chain  load -> |v| -> divide -> two sincospi  is the fixed cost of a state (about half of the time at
// n = 9, nearly all of it at n = 6); evaluated for a pair it is paid once per two states.  Lane 16 h + k returns
// (alpha_k, beta_k) of its state:  RZ RY |0> = alpha |0> + beta |1>,  alpha = c e^{-i phi/2},  beta = s e^{+i phi/2}.
template <int N>
__device__ __forceinline__ void warp_gates2(const double* __restrict__ vA, const double* __restrict__ vB, int vec_len,
                                            int layers, int lane, double2& al, double2& be, double& nrm) {
    const int k = lane & 15;
    const double* __restrict__ v = (lane & 16) ? vB : vA;
    double part = 0.0;
    for (int i = k; i < vec_len; i += 16) part = fma(v[i], v[i], part);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(FULL_MASK, part, o);   // within the 16-lane half
    nrm = sqrt(part);
    const SvaGate g = sva_gate(k < N ? sva_angle(v, vec_len, nrm, N, layers, 0, k) : 0.0);
    al = make_double2(g.c * g.cp, -g.c * g.sp);
    be = make_double2(g.s * g.cp, g.s * g.sp);
}

// First block on |0..0> from the gates held by lanes base .. base + N - 1: the state is their tensor product.
template <int N>
__device__ __forceinline__ void warp_first_block(double2 al, double2 be, int base, int lane, double2 (&st)[1 << (N - 5)]) {
    constexpr int M = N - 5;
    double2 f = make_double2(1.0, 0.0);                    // this lane's factor: the five qubits across lanes
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double2 a = shfl_d2(al, base + M + j), b = shfl_d2(be, base + M + j);
        f = cmul(((lane >> j) & 1) ? b : a, f);
    }
    st[0] = f;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const double2 a = shfl_d2(al, base + k), b = shfl_d2(be, base + k);
#pragma unroll
        for (int x = 0; x < (1 << k); ++x) {
            st[x | (1 << k)] = cmul(b, st[x]);
            st[x] = cmul(a, st[x]);
        }
    }
    warp_cx<N>(st, lane);
}

// Blocks 1 .. layers-1 (dense): lane k evaluates qubit k's gate of the block.
template <int N>
__device__ __forceinline__ void warp_more_blocks(const double* __restrict__ v, int vec_len, double nrm, int layers, int lane,
                                                 double2 (&st)[1 << (N - 5)]) {
    constexpr int M = N - 5, R = 1 << M;
    for (int layer = 1; layer < layers; ++layer) {
        const SvaGate g = sva_gate(lane < N ? sva_angle(v, vec_len, nrm, N, layers, layer, lane) : 0.0);
#pragma unroll
        for (int k = 0; k < M; ++k) {                      // qubits inside a lane
            SvaGate gk;
            gk.c = shfl_d(g.c, k); gk.s = shfl_d(g.s, k); gk.cp = shfl_d(g.cp, k); gk.sp = shfl_d(g.sp, k);
#pragma unroll
            for (int x = 0; x < R; ++x)
                if (!((x >> k) & 1)) sva_apply(gk, st[x], st[x | (1 << k)]);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {                      // qubits across lanes: exchange with lane ^ (1 << j)
            SvaGate gk;
            gk.c = shfl_d(g.c, M + j); gk.s = shfl_d(g.s, M + j); gk.cp = shfl_d(g.cp, M + j); gk.sp = shfl_d(g.sp, M + j);
            const bool hi = (lane >> j) & 1;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double2 other;
                other.x = __shfl_xor_sync(FULL_MASK, st[r].x, 1 << j);
                other.y = __shfl_xor_sync(FULL_MASK, st[r].y, 1 << j);
                double2 a0 = hi ? other : st[r], a1 = hi ? st[r] : other;
                sva_apply(gk, a0, a1);
                st[r] = hi ? a1 : a0;
            }
        }
        warp_cx<N>(st, lane);
    }
}

template <int N, bool QUERY, bool MULTI>
__global__ void __launch_bounds__(SVA_WARP_THREADS) sva_warp_kernel(const SvaParams p) {
    constexpr int R = 1 << (N - 5);
    const int lane = threadIdx.x & 31;
    // the shuffle from lane 0 tells the compiler that the warp index -- and with it the loop below -- is warp-uniform
    // (otherwise every shuffle in the loop is wrapped in a WARPSYNC / ENDCOLLECTIVE pair)
    const int warp_in_cta = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = QUERY ? (int64_t)p.nq : p.nd;
    const double* __restrict__ vecs = QUERY ? p.qvec : p.dvec;
    if (QUERY) griddep_launch_dependents();
    bool waited = false;
    // PAIR: states 2 jj and 2 jj + 1 share one gate pass.  Measured (1000 x 100 pairs): n = 6 68 -> 53 us, n = 8 91 -> 77,
    // n = 9 128 -> 111; at n = 10 and in the multi-block kernels the extra registers cost more than the pass saves
    // (214 -> 225 us; n = 9, 3 blocks: 1.34 -> 1.58 ms), so those take one state per pass.
    constexpr bool PAIR = !MULTI && N <= 9;
    constexpr int PER = PAIR ? 2 : 1;
    for (int64_t jj = warp0; PER * jj < total; jj += nwarps) {
        const int64_t jA = PER * jj;
        const bool hasB = PAIR && jA + 1 < total;
        const int64_t jB = hasB ? jA + 1 : jA;
        double2 al, be;
        double nrm;
        warp_gates2<N>(vecs + jA * p.vec_len, vecs + jB * p.vec_len, p.vec_len, p.layers, lane, al, be, nrm);
#pragma unroll
        for (int h = 0; h < PER; ++h) {
            if (h == 1 && !hasB) break;
            const int64_t j = h ? jB : jA;
            double2 st[R];
            warp_first_block<N>(al, be, 16 * h, lane, st);
            if (MULTI) warp_more_blocks<N>(vecs + j * p.vec_len, p.vec_len, shfl_d(nrm, 16 * h), p.layers, lane, st);
            if (QUERY) {
#pragma unroll
                for (int r = 0; r < R; ++r) p.qstate[(j * R + r) * 32 + lane] = st[r];
                continue;
            }
            const int64_t qi = p.doc_query ? (int64_t)p.doc_query[j] : j / p.docs_per_query;
            if (qi < 0 || qi >= p.nq) {                        // a doc_query entry that names no query: never an OOB read
                if (lane == 0) p.out[j] = sva_nan();
                continue;
            }
            if (!waited) { griddep_wait(); waited = true; }    // the query states are pass 1's output
            const double2* __restrict__ q = p.qstate + qi * R * 32 + lane;
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 d = st[r], a = q[r * 32];
                re = fma(d.x, a.x, fma(d.y, a.y, re));
                im = fma(d.x, a.y, fma(-d.y, a.x, im));
            }
            re = warp_sum(re);
            im = warp_sum(im);
            if (lane == 0) p.out[j] = re * re + im * im;
        }
    }
}

// ===========================================================================
// host side
// ===========================================================================
// Stream-ordered scratch: one pool per device that keeps its memory between calls (the default pool would hand it
// back to the driver at every synchronisation).
static int sva_pool(cudaMemPool_t* out) {
    static cudaMemPool_t pools[64] = {};
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    QRAG_CUDA_CHECK(cudaGetDevice(&dev));
    QRAG_REQUIRE(dev >= 0 && dev < 64, QRAG_ERR_UNSUPPORTED, "device ordinal %d", dev);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        QRAG_CUDA_CHECK(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = ~0ull;
        QRAG_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pools[dev] = pool;
    }
    *out = pools[dev];
    return QRAG_OK;
}

template <typename K>
static int sva_launch(K kern, const SvaParams& p, int64_t grid, int threads, bool chained, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // pass 2 may start while pass 1 runs
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = chained ? 1 : 0;
    QRAG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    return QRAG_OK;
}

template <typename K>
static int sva_warp_grid(K kern, int64_t states, int per_pass, int64_t* grid) {
    const DeviceProps& dp = device_props();
    static int per_sm_cached = 0;                          // one per instantiation of this template (= per kernel)
    int per_sm = per_sm_cached;
    if (per_sm == 0) {
        QRAG_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SVA_WARP_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        per_sm_cached = per_sm;
    }
    const int64_t cap = (int64_t)dp.sm_count * per_sm;     // every CTA resident; warps stride over the states
    const int64_t need = ceil_div(ceil_div(states, per_pass), SVA_WARP_THREADS / 32);   // a warp takes per_pass states per step
    *grid = need < cap ? need : cap;
    return QRAG_OK;
}

template <int N, bool MULTI>
static int sva_run(const SvaParams& p, cudaStream_t st) {
    if constexpr (N <= 5) {
        int rc = sva_launch(sva_thread_kernel<N, true, MULTI>, p, ceil_div(p.nq, 128), 128, false, st);
        if (rc) return rc;
        return sva_launch(sva_thread_kernel<N, false, MULTI>, p, ceil_div(p.nd, 128), 128, true, st);
    } else {
        int64_t gq = 1, gd = 1;
        constexpr int per_pass = (!MULTI && N <= 9) ? 2 : 1;             // = PAIR in sva_warp_kernel
        int rc = sva_warp_grid(sva_warp_kernel<N, true, MULTI>, p.nq, per_pass, &gq);
        if (!rc) rc = sva_warp_grid(sva_warp_kernel<N, false, MULTI>, p.nd, per_pass, &gd);
        if (!rc) rc = sva_launch(sva_warp_kernel<N, true, MULTI>, p, gq, SVA_WARP_THREADS, false, st);
        if (rc) return rc;
        return sva_launch(sva_warp_kernel<N, false, MULTI>, p, gd, SVA_WARP_THREADS, true, st);
    }
}

template <int N>
static int sva_run_n(const SvaParams& p, cudaStream_t st) {
    return p.layers > 1 ? sva_run<N, true>(p, st) : sva_run<N, false>(p, st);
}

// n_qubits in [1, 10]; arguments already validated by qrag_sv_fidelity_angle (sv_kernels.cu)
int sv_angle_registers(const double* qvec, int nq, const double* dvec, int64_t nd, const int32_t* doc_query,
                       int64_t docs_per_query, int vec_len, int n_qubits, int layers, double* out, cudaStream_t st) {
    cudaMemPool_t pool;
    int rc = sva_pool(&pool);
    if (rc) return rc;
    SvaParams p{qvec, dvec, doc_query, nd, docs_per_query, nq, vec_len, layers, nullptr, out};
    const size_t bytes = (size_t)nq * ((size_t)1 << n_qubits) * sizeof(double2);
    void* scratch = nullptr;
    QRAG_CUDA_CHECK(cudaMallocFromPoolAsync(&scratch, bytes, pool, st));
    p.qstate = static_cast<double2*>(scratch);
    switch (n_qubits) {
        case 1: rc = sva_run_n<1>(p, st); break;
        case 2: rc = sva_run_n<2>(p, st); break;
        case 3: rc = sva_run_n<3>(p, st); break;
        case 4: rc = sva_run_n<4>(p, st); break;
        case 5: rc = sva_run_n<5>(p, st); break;
        case 6: rc = sva_run_n<6>(p, st); break;
        case 7: rc = sva_run_n<7>(p, st); break;
        case 8: rc = sva_run_n<8>(p, st); break;
        case 9: rc = sva_run_n<9>(p, st); break;
        case 10: rc = sva_run_n<10>(p, st); break;
        default: rc = set_error(QRAG_ERR_UNSUPPORTED, "register statevector path: n_qubits=%d", n_qubits);
    }
    const cudaError_t fe = cudaFreeAsync(scratch, st);
    if (rc) return rc;
    QRAG_REQUIRE(fe == cudaSuccess, QRAG_ERR_CUDA, "cudaFreeAsync failed: %s", cudaGetErrorString(fe));
    QRAG_LAUNCH_CHECK("sva kernels");
    return QRAG_OK;
}

}  // namespace qrag
