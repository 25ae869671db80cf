// K1: exact statevector simulation of the reference fidelity circuit, batched.
//
// Reference being replaced (one pair at a time, two Qiskit execute() calls per
// pair): QuantumReranker._quantum_similarity / _vector_to_circuit,
// /root/reference/src/reranker/quantum.py:108-167.
//
// complex128 throughout.  n_qubits <= 10 runs with the states in registers (sv_angle.cu); here:
//  * sv_cta_kernel (n_qubits = 11, 12, and the fall-back of the feature map): the 2^n amplitudes are staged in shared memory
//    (double2, XOR-swizzled so that all radix-8 passes are bank-conflict free); each
//    pass applies up to three qubits' RY/RZ in registers; the pass that finishes a
//    layer scatters through the CX-chain permutation into the other buffer.
//    Also serves the amplitude-encoded feature map (fp32 rows, `layers` >= 1).
//
// Gate semantics (Qiskit, little-endian): RY(t) = [[c,-s],[s,c]], c = cos t/2;
// RZ(p) = diag(e^{-ip/2}, e^{+ip/2}); CX(i,i+1) for i = 0..n-2 maps basis x to
// y with y_k = x_0 ^ ... ^ x_k (prefix xor).
#include "common.cuh"
#include "sort.cuh"

namespace qrag {

constexpr double kPi = 3.141592653589793238462643383279502884;

struct GateParams { double c, s, cp, sp; };   // RY half-angle cos/sin, RZ half-angle cos/sin

__device__ __forceinline__ GateParams gate_params(double a) {
    // quantum.py:160-161: ry(a*pi), rz(a*pi/2)
    GateParams g;
    const double theta = a * kPi;
    const double phi = a * kPi / 2.0;
    sincos(theta / 2.0, &g.s, &g.c);
    sincos(0.5 * phi, &g.sp, &g.cp);
    return g;
}

// Apply RZ(phi) RY(theta) to the amplitude pair (a0 = bit clear, a1 = bit set).
__device__ __forceinline__ void apply_gate(const GateParams& g, double2& a0, double2& a1) {
    const double t0r = g.c * a0.x - g.s * a1.x, t0i = g.c * a0.y - g.s * a1.y;
    const double t1r = g.s * a0.x + g.c * a1.x, t1i = g.s * a0.y + g.c * a1.y;
    a0.x = t0r * g.cp + t0i * g.sp;  a0.y = t0i * g.cp - t0r * g.sp;   // * e^{-i phi/2}
    a1.x = t1r * g.cp - t1i * g.sp;  a1.y = t1i * g.cp + t1r * g.sp;   // * e^{+i phi/2}
}

__device__ __forceinline__ int prefix_xor(int x) {
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8;
    return x;
}

// ===========================================================================
// CTA path, n <= 12
// ===========================================================================
struct SvCtaParams {
    // ANGLE mode
    const double* qvec; const double* dvec; const int32_t* doc_query;
    int64_t nd, docs_per_query; int vec_len;
    // FMAP mode
    const float* Q; const float* cand; const float* X; const int64_t* idx; int64_t N; int64_t C; int D;
    int nq, n, layers;
    int64_t docs_per_cta;
    double* out; float* out32;
};

__device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }

// block-wide sum, result broadcast to all threads; `red` has >= 32 doubles.
__device__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect red[] from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// One radix-8 pass over qubits [k0, k0+g).  If `nxt` is non-null this pass ends the
// layer: amplitudes are scattered through the CX-chain permutation into `nxt`.
__device__ void apply_pass(double2* __restrict__ cur, double2* __restrict__ nxt, const GateParams* __restrict__ par,
                           int n, int k0, int g) {
    const int dim = 1 << n, ngroups = dim >> g, na = 1 << g;
    const int lowmask = (1 << k0) - 1;
    for (int t = threadIdx.x; t < ngroups; t += blockDim.x) {
        const int base = ((t & ~lowmask) << g) | (t & lowmask);
        double2 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < na) a[j] = cur[swz(base | (j << k0))];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b < g) {
                const GateParams gp = par[k0 + b];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (!((j >> b) & 1) && (j | (1 << b)) < na) apply_gate(gp, a[j], a[j | (1 << b)]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < na) {
                const int x = base | (j << k0);
                if (nxt) nxt[swz(prefix_xor(x) & (dim - 1))] = a[j];
                else     cur[swz(x)] = a[j];
            }
        }
    }
    __syncthreads();
}

// Runs `layers` blocks on the state in buf[which]; returns the buffer index holding the result.
__device__ int run_layers(double2* buf0, double2* buf1, int which, const GateParams* par, int n, int layers) {
    for (int layer = 0; layer < layers; ++layer) {
        double2* cur = which ? buf1 : buf0;
        double2* nxt = which ? buf0 : buf1;
        for (int k0 = 0; k0 < n; k0 += 3) {
            const int g = (n - k0) < 3 ? (n - k0) : 3;
            const bool last = (k0 + 3 >= n);
            apply_pass(cur, last ? nxt : nullptr, par + layer * n, n, k0, g);
        }
        which ^= 1;
    }
    return which;
}

// Builds one state; returns buffer index.  MODE 0: angle vector (fp64), start |0..0>.
// MODE 1: fp32 row, amplitude-encoded start, angles from the normalised row.
template <int MODE>
__device__ int build_state(const void* src, int len, double2* buf0, double2* buf1, GateParams* par, double* red,
                           int n, int layers, bool* is_zero) {
    const int dim = 1 << n;
    double part = 0.0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const double x = MODE == 0 ? static_cast<const double*>(src)[i] : (double)static_cast<const float*>(src)[i];
        part = fma(x, x, part);
    }
    const double nrm = sqrt(block_sum(part, red));
    *is_zero = (nrm == 0.0);
    const int limit = len < n ? len : n;
    for (int t = threadIdx.x; t < layers * n; t += blockDim.x) {
        const int k = t % n;
        GateParams g;
        if (len == 0 || (MODE == 0 && layers == 1 && k >= limit)) {
            g.c = 1.0; g.s = 0.0; g.cp = 1.0; g.sp = 0.0;              // qubit not rotated (quantum.py:158)
        } else {
            const int comp = t % len;
            const double x = MODE == 0 ? static_cast<const double*>(src)[comp]
                                       : (double)static_cast<const float*>(src)[comp];
            g = gate_params(nrm > 0.0 ? x / nrm : x);
        }
        par[t] = g;
    }
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (MODE == 0) {
            if (i == 0) v.x = 1.0;
        } else if (i < len && nrm > 0.0) {
            v.x = (double)static_cast<const float*>(src)[i] / nrm;
        }
        buf0[swz(i)] = v;
    }
    __syncthreads();
    return run_layers(buf0, buf1, 0, par, n, layers);
}

template <int MODE>
__global__ void sv_cta_kernel(const SvCtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, dim = 1 << n;
    double2* qstate = reinterpret_cast<double2*>(smem_raw);
    double2* buf0 = qstate + dim;
    double2* buf1 = buf0 + dim;
    GateParams* par = reinterpret_cast<GateParams*>(buf1 + dim);
    double* red = reinterpret_cast<double*>(par + p.layers * n);

    const int64_t total = MODE == 0 ? p.nd : (int64_t)p.nq * p.C;
    const int64_t per_query = MODE == 0 ? p.docs_per_query : p.C;
    const int len = MODE == 0 ? p.vec_len : p.D;
    const int64_t j0 = (int64_t)blockIdx.x * p.docs_per_cta;
    int64_t j1 = j0 + p.docs_per_cta;
    if (j1 > total) j1 = total;
    int64_t cur_q = -1;
    bool q_zero = false;

    for (int64_t j = j0; j < j1; ++j) {
        const int64_t qi = (MODE == 0 && p.doc_query) ? (int64_t)p.doc_query[j] : j / per_query;
        if (qi != cur_q) {
            const void* qsrc = MODE == 0 ? (const void*)(p.qvec + qi * len) : (const void*)(p.Q + qi * len);
            bool z;
            const int w = build_state<MODE>(qsrc, len, buf0, buf1, par, red, n, p.layers, &z);
            const double2* res = w ? buf1 : buf0;
            for (int i = threadIdx.x; i < dim; i += blockDim.x) qstate[i] = res[i];
            __syncthreads();
            q_zero = z;
            cur_q = qi;
        }
        const void* dsrc;
        bool missing = false;
        if (MODE == 0) {
            dsrc = p.dvec + j * len;
        } else if (p.cand) {
            dsrc = p.cand + (size_t)j * len;
        } else {
            const int64_t id = p.idx[j];
            missing = id < 0 || id >= p.N;
            dsrc = p.X + (size_t)(missing ? 0 : id) * len;
        }
        bool d_zero;
        const int w = build_state<MODE>(dsrc, len, buf0, buf1, par, red, n, p.layers, &d_zero);
        const double2* ds = w ? buf1 : buf0;
        double re = 0.0, im = 0.0;
        for (int i = threadIdx.x; i < dim; i += blockDim.x) {
            const double2 d = ds[i], q = qstate[i];
            re += d.x * q.x + d.y * q.y;
            im += d.x * q.y - d.y * q.x;
        }
        re = block_sum(re, red);
        im = block_sum(im, red);
        if (threadIdx.x == 0) {
            double f = re * re + im * im;
            if (MODE == 1 && (q_zero || d_zero)) f = 0.0;
            if (missing) f = -pos_inf();
            p.out[j] = f;
            if (p.out32) p.out32[j] = (float)f;
        }
        __syncthreads();
    }
}

static size_t sv_cta_smem(int n, int layers) {
    return (size_t)3 * ((size_t)1 << n) * sizeof(double2) + (size_t)layers * n * sizeof(GateParams) + 32 * sizeof(double);
}

template <int MODE>
static int launch_sv_cta(SvCtaParams p, cudaStream_t st) {
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    const size_t smem = sv_cta_smem(p.n, p.layers);
    QRAG_REQUIRE(smem <= (size_t)dp.max_smem_optin, QRAG_ERR_UNSUPPORTED,
                 "statevector kernel: n_qubits=%d layers=%d needs %zu B shared memory", p.n, p.layers, smem);
    const int64_t total = MODE == 0 ? p.nd : (int64_t)p.nq * p.C;
    int threads = (1 << p.n) / 8;
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    int64_t per_cta = total / ((int64_t)dp.sm_count * 16);
    if (per_cta < 1) per_cta = 1;
    if (per_cta > 32) per_cta = 32;
    p.docs_per_cta = per_cta;
    const int64_t grid = ceil_div(total, per_cta);
    QRAG_REQUIRE(grid <= 0x7fffffff, QRAG_ERR_UNSUPPORTED, "too many documents for one launch");
    auto kern = sv_cta_kernel<MODE>;
    if (smem > 48 * 1024)
        QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, st>>>(p);
    QRAG_LAUNCH_CHECK("sv_cta_kernel");
    return QRAG_OK;
}

int sv_angle_registers(const double* qvec, int nq, const double* dvec, int64_t nd, const int32_t* doc_query,
                       int64_t docs_per_query, int vec_len, int n_qubits, int layers, double* out, cudaStream_t st);  // sv_angle.cu

int fmap_warp_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                  int n_qubits, int layers, double* out64, float* out32, cudaStream_t st, bool* handled);  // fmap_warp.cu

int fmap_fidelity(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                  int n_qubits, int layers, double* out64, float* out32, cudaStream_t st) {
    QRAG_REQUIRE(n_qubits <= QRAG_MAX_QUBITS, QRAG_ERR_UNSUPPORTED, "feature map: n_qubits=%d > %d", n_qubits,
                 QRAG_MAX_QUBITS);
    QRAG_REQUIRE(layers <= 64, QRAG_ERR_UNSUPPORTED, "feature map: layers=%d > 64", layers);
    if (fmap_kernel_mode() == QRAG_FMAP_AUTO) {
        bool handled = false;
        const int rc = fmap_warp_try(Q, nq, cand, X, N, idx, C, D, n_qubits, layers, out64, out32, st, &handled);
        if (rc || handled) return rc;
    }
    SvCtaParams p{};
    p.Q = Q; p.cand = cand; p.X = X; p.idx = idx; p.N = N; p.C = C; p.D = D; p.nq = nq; p.n = n_qubits; p.layers = layers;
    p.out = out64; p.out32 = out32;
    return launch_sv_cta<1>(p, st);
}

// ===========================================================================
// Legacy MT19937 text-hash embedding, quantum.py:169-185
// ===========================================================================
__global__ void mock_embedding_kernel(const uint32_t* __restrict__ seeds, int64_t n, int n_qubits,
                                      double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int m = 4 * n_qubits;              // 32-bit outputs needed (two per double)
    constexpr int MAXM = 4 * QRAG_MAX_QUBITS;
    uint32_t lo[MAXM + 1];                   // mt[0 .. m]
    uint32_t hi[MAXM];                       // mt[397 .. 397+m-1]
    uint32_t x = seeds[t];                   // init_genrand (np.random.seed with a 32-bit int)
    lo[0] = x;
    for (int i = 1; i < 397 + m; ++i) {
        x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
        if (i <= m) lo[i] = x;
        if (i >= 397) hi[i - 397] = x;
    }
    double* o = out + t * (2 * n_qubits);
    double n2 = 0.0;
    uint32_t prev = 0;
    for (int i = 0; i < m; ++i) {
        uint32_t y = (lo[i] & 0x80000000u) | (lo[i + 1] & 0x7fffffffu);
        y = hi[i] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        if (i & 1) {
            const double d = ((double)(prev >> 5) * 67108864.0 + (double)(y >> 6)) / 9007199254740992.0;
            o[i >> 1] = d;
            n2 = fma(d, d, n2);
        }
        prev = y;
    }
    const double nrm = sqrt(n2);
    for (int i = 0; i < 2 * n_qubits; ++i) o[i] = o[i] / nrm;
}

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_sv_fidelity_angle(const double* qvec, int nq, const double* dvec, int64_t nd,
                                      const int32_t* doc_query, int64_t docs_per_query, int vec_len, int n_qubits,
                                      int layers, double* out_scores, void* stream) {
    QRAG_REQUIRE(qvec && dvec && out_scores, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(nq >= 1 && nd >= 0, QRAG_ERR_INVALID, "bad sizes nq=%d nd=%lld", nq, (long long)nd);
    QRAG_REQUIRE(vec_len >= 1, QRAG_ERR_INVALID, "vec_len=%d", vec_len);
    QRAG_REQUIRE(n_qubits >= 1 && n_qubits <= QRAG_MAX_QUBITS, QRAG_ERR_UNSUPPORTED,
                 "n_qubits=%d outside [1, %d]", n_qubits, QRAG_MAX_QUBITS);
    QRAG_REQUIRE(layers >= 1 && layers <= 64, QRAG_ERR_INVALID, "layers=%d outside [1, 64]", layers);
    QRAG_REQUIRE(doc_query != nullptr || docs_per_query >= 1, QRAG_ERR_INVALID,
                 "need doc_query or docs_per_query >= 1");
    QRAG_REQUIRE(doc_query != nullptr || nd <= (int64_t)nq * docs_per_query, QRAG_ERR_INVALID,
                 "nd=%lld exceeds nq*docs_per_query", (long long)nd);
    if (nd == 0) return QRAG_OK;
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_qubits <= 10) {                                   // states in registers (sv_angle.cu)
        QRAG_REQUIRE(ceil_div(nd, 128) <= 0x7fffffff, QRAG_ERR_UNSUPPORTED, "too many documents for one launch");
        return sv_angle_registers(qvec, nq, dvec, nd, doc_query, doc_query ? 1 : docs_per_query, vec_len, n_qubits,
                                  layers, out_scores, st);
    }
    SvCtaParams p{};
    p.qvec = qvec; p.dvec = dvec; p.doc_query = doc_query; p.nd = nd;
    p.docs_per_query = doc_query ? 1 : docs_per_query;
    p.vec_len = vec_len; p.nq = nq; p.n = n_qubits; p.layers = layers; p.out = out_scores;
    return launch_sv_cta<0>(p, st);
}

extern "C" int qrag_mock_embedding(const uint32_t* seeds, int64_t n, int n_qubits, double* out, void* stream) {
    QRAG_REQUIRE(seeds && out, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(n >= 0, QRAG_ERR_INVALID, "n=%lld", (long long)n);
    QRAG_REQUIRE(n_qubits >= 1 && n_qubits <= QRAG_MAX_QUBITS, QRAG_ERR_UNSUPPORTED,
                 "n_qubits=%d outside [1, %d]", n_qubits, QRAG_MAX_QUBITS);
    if (n == 0) return QRAG_OK;
    QRAG_REQUIRE(device_props().ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    const int64_t grid = ceil_div(n, 128);
    mock_embedding_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(seeds, n, n_qubits, out);
    QRAG_LAUNCH_CHECK("mock_embedding_kernel");
    return QRAG_OK;
}
