// ldeq_api.cu -- handle, right-hand-side objects and the GOKU solve entry points of libldeq.so.
//
// ldeq_solve_fwd / ldeq_solve_bwd are the drop-in for the body of
// diffeq_layer(::Decoder{<:GOKU}, l, t) (reference src/models/GOKU.jl:98-130) and its pullback.
#include <cstdio>
#include <cstring>

#include "ldeq_internal.h"
#include "ldeq_rhs.cuh"
#include "ldeq_tsit5.cuh"

namespace ldeq {

int set_err(ldeq_handle* h, int code, const char* what, cudaError_t ce) {
    if (h) {
        h->err = what ? what : "";
        if (ce != cudaSuccess) {
            h->err += ": ";
            h->err += cudaGetErrorString(ce);
        }
    }
    return code;
}

int upload_tgrid(ldeq_handle* h, const double* t_host, int T, cudaStream_t s) {
    if ((size_t)T > h->d_tgrid_cap) {
        // stream-ordered free so that kernels still reading the old grid finish first
        if (h->d_tgrid) LDEQ_CUDA(cudaFreeAsync(h->d_tgrid, s));
        h->d_tgrid = nullptr;
        h->d_tgrid_cap = 0;
        size_t cap = (size_t)T < 256 ? 256 : (size_t)T;
        LDEQ_CUDA(cudaMallocAsync((void**)&h->d_tgrid, cap * sizeof(double), s));
        h->d_tgrid_cap = cap;
        h->h_tgrid.clear();
    }
    if (h->h_tgrid.size() != (size_t)T || std::memcmp(h->h_tgrid.data(), t_host, (size_t)T * sizeof(double)) != 0) {
        h->h_tgrid.assign(t_host, t_host + T);
        // source is pageable host memory: the runtime stages it before returning
        LDEQ_CUDA(cudaMemcpyAsync(h->d_tgrid, t_host, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    return LDEQ_OK;
}

int ensure_scratch(ldeq_handle* h, int slot, size_t bytes) {
    if (bytes <= h->scratch_cap[slot]) return LDEQ_OK;
    if (h->scratch[slot]) {
        LDEQ_CUDA(cudaDeviceSynchronize());
        LDEQ_CUDA(cudaFree(h->scratch[slot]));
        h->scratch[slot] = nullptr;
        h->scratch_cap[slot] = 0;
    }
    size_t cap = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&h->scratch[slot], cap);
    if (e != cudaSuccess) return set_err(h, LDEQ_ERR_NOMEM, "cudaMalloc(scratch)", e);
    h->scratch_cap[slot] = cap;
    return LDEQ_OK;
}

KOpts to_kopts(const ldeq_opts* o) {
    KOpts k;
    k.abstol = o->abstol; k.reltol = o->reltol; k.dt = o->dt; k.dtmax = o->dtmax; k.dtmin = o->dtmin;
    k.gamma = o->gamma; k.qmin = o->qmin; k.qmax = o->qmax; k.beta1 = o->beta1; k.beta2 = o->beta2;
    k.qoldinit = o->qoldinit; k.qsteady_min = o->qsteady_min; k.qsteady_max = o->qsteady_max;
    k.maxiters = o->maxiters; k.adaptive = o->adaptive; k.controller_pow = o->controller_pow;
    return k;
}

template <class S, bool FRICTION, bool TAPE>
static cudaError_t launch_fwd(const void* z0, const void* theta, const double* tg, int B, int T, const KOpts& ko,
                              void* traj, int32_t* ret, int32_t* na, int32_t* nr, const ldeq_tape* tape,
                              cudaStream_t s) {
    TapeView<S> tv{nullptr, nullptr, nullptr, nullptr, 0};
    if (TAPE) tv = TapeView<S>{tape->t, tape->dt, (S*)tape->u, tape->overflow, tape->cap};
    const int grid = (B + LDEQ_FWD_THREADS - 1) / LDEQ_FWD_THREADS;
    tsit5_fwd_kernel<PendulumRHS<S, FRICTION>, S, TAPE><<<grid, LDEQ_FWD_THREADS, 0, s>>>(
        (const S*)z0, (const S*)theta, tg, B, T, ko, (S*)traj, ret, na, nr, tv);
    return cudaGetLastError();
}

template <class S, bool FRICTION>
static cudaError_t launch_bwd(const ldeq_tape* tape, const void* dtraj, void* dz0, void* dtheta, cudaStream_t s) {
    TapeView<S> tv{tape->t, tape->dt, (S*)tape->u, tape->overflow, tape->cap};
    const int grid = (tape->B + LDEQ_BWD_THREADS - 1) / LDEQ_BWD_THREADS;
    tsit5_bwd_kernel<PendulumRHS<S, FRICTION>, S><<<grid, LDEQ_BWD_THREADS, 0, s>>>(
        (const S*)tape->theta, tape->tgrid, tape->B, tape->T, (const S*)dtraj, tv, tape->retcode, tape->naccept,
        (S*)dz0, (S*)dtheta);
    return cudaGetLastError();
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace ldeq

using namespace ldeq;

extern "C" {

int ldeq_version(void) { return LDEQ_VERSION; }

void ldeq_opts_default(ldeq_opts* o) {
    std::memset(o, 0, sizeof(*o));
    o->abstol = 1e-6; o->reltol = 1e-3; o->adaptive = 1; o->controller_pow = 0; o->dt = 0.0; o->dtmax = 0.0;
    o->dtmin = 0.0; o->maxiters = 1000000; o->gamma = 0.9; o->qmin = 0.2; o->qmax = 10.0; o->beta1 = 7.0 / 50.0;
    o->beta2 = 2.0 / 25.0; o->qoldinit = 1e-4; o->qsteady_min = 1.0; o->qsteady_max = 1.0; o->tape_steps = 0;
    o->norm_mode = LDEQ_NORM_GLOBAL; o->mlp_math = LDEQ_MLP_MATH_FP32;
}

int ldeq_create(ldeq_handle** out, int device) {
    if (!out) return LDEQ_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return LDEQ_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return LDEQ_ERR_CUDA;
    ldeq_handle* h = new ldeq_handle();
    h->device = device;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    // keep stream-ordered allocations (tapes) in the pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = h;
    return LDEQ_OK;
}

void ldeq_destroy(ldeq_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->d_tgrid) cudaFree(h->d_tgrid);
    for (int i = 0; i < 4; ++i)
        if (h->scratch[i]) cudaFree(h->scratch[i]);
    if (h->d_partials) cudaFree(h->d_partials);
    if (h->d_counter) cudaFree(h->d_counter);
    delete h;
}

const char* ldeq_last_error(const ldeq_handle* h) { return h ? h->err.c_str() : "null handle"; }
int64_t ldeq_launch_count(const ldeq_handle* h) { return h ? h->launches : 0; }

int ldeq_rhs_builtin(ldeq_handle* h, int kind, ldeq_rhs** out) {
    if (!h || !out) return LDEQ_ERR_INVALID;
    if (kind != LDEQ_RHS_PENDULUM && kind != LDEQ_RHS_PENDULUM_FRICTION)
        return set_err(h, LDEQ_ERR_INVALID, "unknown built-in rhs kind");
    ldeq_rhs* r = new ldeq_rhs();
    r->kind = kind;
    r->z_dim = 2;
    r->p_dim = 1;
    *out = r;
    return LDEQ_OK;
}

int ldeq_rhs_dims(const ldeq_rhs* rhs, int* z_dim, int* p_dim) {
    if (!rhs) return LDEQ_ERR_INVALID;
    if (z_dim) *z_dim = rhs->z_dim;
    if (p_dim) *p_dim = rhs->p_dim;
    return LDEQ_OK;
}

int ldeq_solve_fwd(ldeq_handle* h, const ldeq_rhs* rhs, int dtype, const void* z0, const void* theta,
                   const double* t_host, int B, int T, const ldeq_opts* opts, void* traj_out, int32_t* retcode,
                   int32_t* naccept, int32_t* nreject, ldeq_tape** tape_out, ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (tape_out) *tape_out = nullptr;
    if (!rhs || !z0 || !theta || !t_host || !traj_out || !opts) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    if (B < 0 || T < 1) return set_err(h, LDEQ_ERR_INVALID, "B must be >= 0 and T >= 1");
    if (dtype != LDEQ_F32 && dtype != LDEQ_F64) return set_err(h, LDEQ_ERR_INVALID, "dtype");
    if (!opts->adaptive && !(opts->dt > 0.0)) return set_err(h, LDEQ_ERR_INVALID, "adaptive = 0 needs dt > 0");
    for (int k = 1; k < T; ++k)
        if (!(t_host[k] > t_host[k - 1])) return set_err(h, LDEQ_ERR_INVALID, "t must be strictly increasing");
    if (rhs->kind < 0) return set_err(h, LDEQ_ERR_UNSUPPORTED, "user rhs goes through ldeq_rhs_from_source path");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    if (B == 0) return LDEQ_OK;
    int rc = upload_tgrid(h, t_host, T, s);
    if (rc) return rc;
    KOpts ko = to_kopts(opts);
    const size_t es = dtype == LDEQ_F32 ? 4 : 8;
    const int ZD = rhs->z_dim, PD = rhs->p_dim;

    ldeq_tape* tape = nullptr;
    if (tape_out) {
        tape = new ldeq_tape();
        tape->dtype = dtype; tape->rhs_kind = rhs->kind; tape->rhs = rhs; tape->B = B; tape->T = T;
        tape->z_dim = ZD; tape->p_dim = PD;
        long long cap = opts->tape_steps;
        if (cap <= 0) {
            if (opts->adaptive) {
                cap = 2LL * T > 64 ? 2LL * T : 64;
            } else {
                cap = (long long)((t_host[T - 1] - t_host[0]) / opts->dt) + 3;
            }
        }
        if (cap > opts->maxiters) cap = opts->maxiters;
        if (cap > (1 << 24)) cap = 1 << 24;
        tape->cap = (int)cap;
        const size_t nB = (size_t)B, c = (size_t)cap;
        size_t off = 0, o_t = off;   off += align_up(c * nB * 8);
        size_t o_dt = off;           off += align_up(c * nB * 8);
        size_t o_u = off;            off += align_up(c * nB * ZD * es);
        size_t o_th = off;           off += align_up(nB * PD * es);
        size_t o_tg = off;           off += align_up((size_t)T * 8);
        size_t o_ret = off;          off += align_up(nB * 4);
        size_t o_na = off;           off += align_up(nB * 4);
        size_t o_nr = off;           off += align_up(nB * 4);
        size_t o_ov = off;           off += 256;
        cudaError_t e = cudaMallocAsync(&tape->base, off, s);
        if (e != cudaSuccess) {
            delete tape;
            return set_err(h, LDEQ_ERR_NOMEM, "cudaMallocAsync(tape)", e);
        }
        char* bp = (char*)tape->base;
        tape->t = (double*)(bp + o_t); tape->dt = (double*)(bp + o_dt); tape->u = bp + o_u;
        tape->theta = bp + o_th; tape->tgrid = (double*)(bp + o_tg); tape->retcode = (int32_t*)(bp + o_ret);
        tape->naccept = (int32_t*)(bp + o_na); tape->nreject = (int32_t*)(bp + o_nr);
        tape->overflow = (int32_t*)(bp + o_ov);
        cudaMemcpyAsync(tape->theta, theta, nB * PD * es, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(tape->tgrid, h->d_tgrid, (size_t)T * 8, cudaMemcpyDeviceToDevice, s);
        cudaMemsetAsync(tape->overflow, 0, 4, s);
    }
    int32_t* d_ret = tape ? tape->retcode : retcode;
    int32_t* d_na = tape ? tape->naccept : naccept;
    int32_t* d_nr = tape ? tape->nreject : nreject;
    const bool fr = rhs->kind == LDEQ_RHS_PENDULUM_FRICTION;
    cudaError_t e;
#define LDEQ_DISPATCH_FWD(S)                                                                                         \
    (tape ? (fr ? launch_fwd<S, true, true>(z0, theta, h->d_tgrid, B, T, ko, traj_out, d_ret, d_na, d_nr, tape, s)  \
                : launch_fwd<S, false, true>(z0, theta, h->d_tgrid, B, T, ko, traj_out, d_ret, d_na, d_nr, tape, s)) \
          : (fr ? launch_fwd<S, true, false>(z0, theta, h->d_tgrid, B, T, ko, traj_out, d_ret, d_na, d_nr, tape, s) \
                : launch_fwd<S, false, false>(z0, theta, h->d_tgrid, B, T, ko, traj_out, d_ret, d_na, d_nr, tape, s)))
    e = dtype == LDEQ_F32 ? LDEQ_DISPATCH_FWD(float) : LDEQ_DISPATCH_FWD(double);
#undef LDEQ_DISPATCH_FWD
    h->launches += 1;
    if (e != cudaSuccess) {
        if (tape) { cudaFreeAsync(tape->base, s); delete tape; }
        return set_err(h, LDEQ_ERR_CUDA, "tsit5_fwd_kernel launch", e);
    }
    if (tape) {
        if (retcode) cudaMemcpyAsync(retcode, tape->retcode, (size_t)B * 4, cudaMemcpyDeviceToDevice, s);
        if (naccept) cudaMemcpyAsync(naccept, tape->naccept, (size_t)B * 4, cudaMemcpyDeviceToDevice, s);
        if (nreject) cudaMemcpyAsync(nreject, tape->nreject, (size_t)B * 4, cudaMemcpyDeviceToDevice, s);
        *tape_out = tape;
    }
    return LDEQ_OK;
}

int ldeq_solve_bwd(ldeq_handle* h, ldeq_tape* tape, const void* dtraj, void* dz0, void* dtheta, ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (!tape || !dtraj || !dz0 || !dtheta) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    const bool fr = tape->rhs_kind == LDEQ_RHS_PENDULUM_FRICTION;
    cudaError_t e;
    if (tape->dtype == LDEQ_F32)
        e = fr ? launch_bwd<float, true>(tape, dtraj, dz0, dtheta, s) : launch_bwd<float, false>(tape, dtraj, dz0, dtheta, s);
    else
        e = fr ? launch_bwd<double, true>(tape, dtraj, dz0, dtheta, s) : launch_bwd<double, false>(tape, dtraj, dz0, dtheta, s);
    h->launches += 1;
    if (e != cudaSuccess) return set_err(h, LDEQ_ERR_CUDA, "tsit5_bwd_kernel launch", e);
    return LDEQ_OK;
}

int ldeq_tape_overflow(ldeq_handle* h, ldeq_tape* tape, int32_t* count_host, ldeq_stream stream) {
    if (!h || !tape || !count_host) return LDEQ_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaMemcpyAsync(count_host, tape->overflow, 4, cudaMemcpyDeviceToHost, s));
    LDEQ_CUDA(cudaStreamSynchronize(s));
    return LDEQ_OK;
}

void ldeq_tape_free(ldeq_handle* h, ldeq_tape* tape, ldeq_stream stream) {
    if (!tape) return;
    if (h) cudaSetDevice(h->device);
    if (tape->base) cudaFreeAsync(tape->base, (cudaStream_t)stream);
    delete tape;
}

int ldeq_solve_fwd_host(ldeq_handle* h, const ldeq_rhs* rhs, int dtype, const void* z0_host, const void* theta_host,
                        const double* t_host, int B, int T, const ldeq_opts* opts, void* traj_out_host,
                        int32_t* retcode_host, int32_t* naccept_host, int32_t* nreject_host, ldeq_tape** tape_out,
                        ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (!rhs || !z0_host || !theta_host || !traj_out_host) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    if (B <= 0 || T < 1) return set_err(h, LDEQ_ERR_INVALID, "B must be > 0 and T >= 1");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    const size_t es = dtype == LDEQ_F32 ? 4 : 8;
    const size_t nz = (size_t)B * rhs->z_dim * es, np = (size_t)B * rhs->p_dim * es, nt = nz * (size_t)T;
    int rc;
    if ((rc = ensure_scratch(h, 0, align_up(nz) + align_up(np) + 3 * align_up((size_t)B * 4)))) return rc;
    if ((rc = ensure_scratch(h, 1, nt))) return rc;
    char* in = (char*)h->scratch[0];
    void* d_z0 = in;
    void* d_th = in + align_up(nz);
    int32_t* d_ret = (int32_t*)(in + align_up(nz) + align_up(np));
    int32_t* d_na = (int32_t*)((char*)d_ret + align_up((size_t)B * 4));
    int32_t* d_nr = (int32_t*)((char*)d_na + align_up((size_t)B * 4));
    LDEQ_CUDA(cudaMemcpyAsync(d_z0, z0_host, nz, cudaMemcpyHostToDevice, s));
    LDEQ_CUDA(cudaMemcpyAsync(d_th, theta_host, np, cudaMemcpyHostToDevice, s));
    rc = ldeq_solve_fwd(h, rhs, dtype, d_z0, d_th, t_host, B, T, opts, h->scratch[1], d_ret, d_na, d_nr, tape_out, s);
    if (rc) return rc;
    LDEQ_CUDA(cudaMemcpyAsync(traj_out_host, h->scratch[1], nt, cudaMemcpyDeviceToHost, s));
    if (retcode_host) LDEQ_CUDA(cudaMemcpyAsync(retcode_host, d_ret, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (naccept_host) LDEQ_CUDA(cudaMemcpyAsync(naccept_host, d_na, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (nreject_host) LDEQ_CUDA(cudaMemcpyAsync(nreject_host, d_nr, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    LDEQ_CUDA(cudaStreamSynchronize(s));
    return LDEQ_OK;
}

int ldeq_solve_bwd_host(ldeq_handle* h, ldeq_tape* tape, const void* dtraj_host, void* dz0_host, void* dtheta_host,
                        ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (!tape || !dtraj_host || !dz0_host || !dtheta_host) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    const size_t es = tape->dtype == LDEQ_F32 ? 4 : 8;
    const size_t nz = (size_t)tape->B * tape->z_dim * es, np = (size_t)tape->B * tape->p_dim * es,
                 nt = nz * (size_t)tape->T;
    int rc;
    if ((rc = ensure_scratch(h, 2, nt))) return rc;
    if ((rc = ensure_scratch(h, 3, align_up(nz) + align_up(np)))) return rc;
    void* d_dz0 = h->scratch[3];
    void* d_dth = (char*)h->scratch[3] + align_up(nz);
    LDEQ_CUDA(cudaMemcpyAsync(h->scratch[2], dtraj_host, nt, cudaMemcpyHostToDevice, s));
    rc = ldeq_solve_bwd(h, tape, h->scratch[2], d_dz0, d_dth, s);
    if (rc) return rc;
    LDEQ_CUDA(cudaMemcpyAsync(dz0_host, d_dz0, nz, cudaMemcpyDeviceToHost, s));
    LDEQ_CUDA(cudaMemcpyAsync(dtheta_host, d_dth, np, cudaMemcpyDeviceToHost, s));
    LDEQ_CUDA(cudaStreamSynchronize(s));
    return LDEQ_OK;
}

}  // extern "C"
