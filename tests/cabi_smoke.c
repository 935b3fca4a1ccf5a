/* cabi_smoke.c -- a plain C program that drives libldeq.so through include/ldeq.h alone (no Python, no torch): the
 * nearest executable stand-in for the `ccall` sequence of julia/LatentDiffEqB200.jl.
 *
 *   gcc -O1 -I include -I /usr/local/cuda/include tests/cabi_smoke.c -o tests/cabi_smoke \
 *       -L latentdiffeq.jl_b200/lib -lldeq -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,...
 *
 * Exit code 0 = every check passed.  Needs a CUDA device. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ldeq.h"

#define CHECK(cond, ...)                                          \
    do {                                                          \
        if (!(cond)) {                                            \
            fprintf(stderr, "cabi_smoke FAILED at line %d: ", __LINE__); \
            fprintf(stderr, __VA_ARGS__);                         \
            fprintf(stderr, "\n");                                \
            return 1;                                             \
        }                                                         \
    } while (0)
#define LD(call)                                                                            \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        CHECK(rc_ == LDEQ_OK, "%s -> %d (%s)", #call, rc_, ldeq_last_error(h));             \
    } while (0)

static double loss_of(ldeq_handle* h, const ldeq_rhs* rhs, const float* z0, const float* th, const double* t, int B, int T,
                      const ldeq_opts* o, const float* w, float* traj) {
    /* L = sum w .* traj through the HOST entry point */
    if (ldeq_solve_fwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, o, traj, NULL, NULL, NULL, NULL, NULL) != LDEQ_OK) return NAN;
    double L = 0.0;
    for (size_t i = 0; i < (size_t)B * T * 2; ++i) L += (double)w[i] * traj[i];
    return L;
}

int main(void) {
    enum { B = 300, T = 50, Z = 2, P = 1 };
    ldeq_handle* h = NULL;
    CHECK(ldeq_version() == LDEQ_VERSION, "header/library version mismatch: %d vs %d", LDEQ_VERSION, ldeq_version());
    CHECK(ldeq_create(&h, 0) == LDEQ_OK && h, "ldeq_create(0): no usable CUDA device");
    ldeq_rhs* rhs = NULL;
    LD(ldeq_rhs_builtin(h, LDEQ_RHS_PENDULUM_FRICTION, &rhs));
    int zd = 0, pd = 0;
    LD(ldeq_rhs_dims(rhs, &zd, &pd));
    CHECK(zd == Z && pd == P, "rhs dims %d %d", zd, pd);
    ldeq_opts o;
    ldeq_opts_default(&o);
    CHECK(o.abstol == 1e-6 && o.reltol == 1e-3 && o.adaptive == 1 && o.sensealg == LDEQ_SENSE_FORWARD_DUAL && o.solver == LDEQ_SOLVER_TSIT5,
          "defaults");

    static float z0[B * Z], th[B * P], w[T * B * Z], traj[T * B * Z], traj2[T * B * Z], dz0[B * Z], dth[B * P], dz0b[B * Z], dthb[B * P];
    double t[T];
    unsigned s = 12345u;
    for (int b = 0; b < B; ++b) {
        s = s * 1664525u + 1013904223u; z0[b * 2] = ((s >> 8) / 16777216.0f - 0.5f) * 1.0f;
        s = s * 1664525u + 1013904223u; z0[b * 2 + 1] = ((s >> 8) / 16777216.0f - 0.5f) * 2.0f;
        s = s * 1664525u + 1013904223u; th[b] = 1.0f + (s >> 8) / 16777216.0f;
    }
    for (int i = 0; i < T * B * Z; ++i) { s = s * 1664525u + 1013904223u; w[i] = (s >> 8) / 16777216.0f - 0.5f; }
    for (int k = 0; k < T; ++k) t[k] = 0.05 * k;

    /* device-pointer path: upload, solve, pullback (both sensitivity modes), download */
    float *d_z0, *d_th, *d_traj, *d_w, *d_dz0, *d_dth;
    int32_t *d_ret, *d_na;
    CHECK(cudaMalloc((void**)&d_z0, sizeof z0) == cudaSuccess && cudaMalloc((void**)&d_th, sizeof th) == cudaSuccess &&
          cudaMalloc((void**)&d_traj, sizeof traj) == cudaSuccess && cudaMalloc((void**)&d_w, sizeof w) == cudaSuccess &&
          cudaMalloc((void**)&d_dz0, sizeof dz0) == cudaSuccess && cudaMalloc((void**)&d_dth, sizeof dth) == cudaSuccess &&
          cudaMalloc((void**)&d_ret, B * 4) == cudaSuccess && cudaMalloc((void**)&d_na, B * 4) == cudaSuccess, "cudaMalloc");
    cudaMemcpy(d_z0, z0, sizeof z0, cudaMemcpyHostToDevice);
    cudaMemcpy(d_th, th, sizeof th, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, w, sizeof w, cudaMemcpyHostToDevice);
    for (int sense = 0; sense < 2; ++sense) {
        o.sensealg = sense;
        ldeq_tape* tape = NULL;
        LD(ldeq_solve_fwd(h, rhs, LDEQ_F32, d_z0, d_th, t, B, T, &o, d_traj, d_ret, d_na, NULL, &tape, NULL));
        CHECK(tape != NULL, "no tape");
        LD(ldeq_solve_bwd(h, tape, d_w, d_dz0, d_dth, NULL));
        int32_t over = -1;
        LD(ldeq_tape_overflow(h, tape, &over, NULL));
        CHECK(over == 0, "tape overflow %d", over);
        ldeq_tape_free(h, tape, NULL);
        CHECK(cudaDeviceSynchronize() == cudaSuccess, "sync");
        cudaMemcpy(sense ? dz0 : dz0b, d_dz0, sizeof dz0, cudaMemcpyDeviceToHost);
        cudaMemcpy(sense ? dth : dthb, d_dth, sizeof dth, cudaMemcpyDeviceToHost);
    }
    cudaMemcpy(traj, d_traj, sizeof traj, cudaMemcpyDeviceToHost);
    static int32_t ret[B];
    cudaMemcpy(ret, d_ret, sizeof ret, cudaMemcpyDeviceToHost);
    for (int b = 0; b < B; ++b) CHECK(ret[b] == LDEQ_RET_SUCCESS, "retcode[%d] = %d", b, ret[b]);
    for (int b = 0; b < B; ++b) CHECK(traj[b * 2] == z0[b * 2] && traj[b * 2 + 1] == z0[b * 2 + 1], "save point 0 must be u0 itself");
    /* the two sensitivity algorithms agree within the solver tolerance */
    double num = 0, den = 0;
    for (int i = 0; i < B * Z; ++i) { num = fmax(num, fabs(dz0[i] - dz0b[i])); den = fmax(den, fabs(dz0[i])); }
    CHECK(num <= 5e-2 * den, "forward-dual vs adjoint dz0: %g of %g", num, den);

    /* host-buffer path: same numbers, and the combined call */
    o.sensealg = LDEQ_SENSE_FORWARD_DUAL;
    ldeq_tape* tape = NULL;
    LD(ldeq_solve_fwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &o, traj2, NULL, NULL, NULL, &tape, NULL));
    CHECK(memcmp(traj, traj2, sizeof traj) == 0, "host and device paths differ");
    static float gz[B * Z], gp[B * P], gz2[B * Z], gp2[B * P];
    LD(ldeq_solve_bwd_host(h, tape, w, gz, gp, NULL));
    ldeq_tape_free(h, tape, NULL);
    CHECK(memcmp(gz, dz0, sizeof gz) == 0 && memcmp(gp, dth, sizeof gp) == 0, "host pullback differs from the device pullback");
    LD(ldeq_solve_fwd_bwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &o, w, traj2, gz2, gp2, NULL, NULL, NULL, NULL));
    CHECK(memcmp(traj, traj2, sizeof traj) == 0 && memcmp(gz, gz2, sizeof gz) == 0 && memcmp(gp, gp2, sizeof gp) == 0, "combined call differs");

    /* finite differences of L = sum w .* traj in fixed-step mode (a smooth function of the inputs) */
    ldeq_opts of;
    ldeq_opts_default(&of);
    of.adaptive = 0; of.dt = 0.05;
    LD(ldeq_solve_fwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &of, traj2, NULL, NULL, NULL, &tape, NULL));
    LD(ldeq_solve_bwd_host(h, tape, w, gz, gp, NULL));
    ldeq_tape_free(h, tape, NULL);
    const float eps = 2e-3f;
    for (int probe = 0; probe < 3; ++probe) {
        const int b = 7 + 97 * probe;
        float save = th[b];
        th[b] = save + eps; const double Lp = loss_of(h, rhs, z0, th, t, B, T, &of, w, traj2);
        th[b] = save - eps; const double Lm = loss_of(h, rhs, z0, th, t, B, T, &of, w, traj2);
        th[b] = save;
        const double fd = (Lp - Lm) / (2.0 * eps);
        CHECK(fabs(fd - gp[b]) <= 2e-2 * fmax(fabs(fd), 1e-3), "dL/dtheta[%d]: finite difference %g vs pullback %g", b, fd, gp[b]);
    }

    /* the diffeq struct's `solver` field: DP5 / BS3 adaptive, RK4 fixed step -- same problem, so the trajectories agree with
       Tsit5's within the solver tolerance, and every solver's pullback is finite and close to Tsit5's */
    for (int sv = LDEQ_SOLVER_DP5; sv <= LDEQ_SOLVER_RK4; ++sv) {
        ldeq_opts os;
        LD(ldeq_opts_default_solver(&os, sv));
        CHECK(os.solver == sv && os.beta2 > 0.0 && os.sensealg == LDEQ_SENSE_FORWARD_DUAL, "ldeq_opts_default_solver(%d)", sv);
        if (sv == LDEQ_SOLVER_RK4) {
            CHECK(ldeq_solve_fwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &os, traj2, NULL, NULL, NULL, NULL, NULL) == LDEQ_ERR_UNSUPPORTED,
                  "adaptive RK4 must be refused");
            os.adaptive = 0; os.dt = 0.025;
        }
        LD(ldeq_solve_fwd_bwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &os, w, traj2, gz2, gp2, ret, NULL, NULL, NULL));
        double dmax = 0, gmax = 0, gden = 0;
        for (int i = 0; i < T * B * Z; ++i) dmax = fmax(dmax, fabs(traj2[i] - traj[i]));
        for (int i = 0; i < B * Z; ++i) { gmax = fmax(gmax, fabs(gz2[i] - dz0[i])); gden = fmax(gden, fabs(dz0[i])); }
        for (int b = 0; b < B; ++b) CHECK(ret[b] == LDEQ_RET_SUCCESS, "solver %d: retcode[%d] = %d", sv, b, ret[b]);
        CHECK(dmax <= 2e-2 && gmax <= 1e-1 * gden, "solver %d vs Tsit5: trajectories %g, dz0 %g of %g", sv, dmax, gmax, gden);
    }

    /* a trajectory that cannot finish: NaN block, zero gradient, no error code (GOKU.jl:114) */
    o.maxiters = 40;
    th[5] = 1e-4f;
    LD(ldeq_solve_fwd_host(h, rhs, LDEQ_F32, z0, th, t, B, T, &o, traj2, ret, NULL, NULL, &tape, NULL));
    LD(ldeq_solve_bwd_host(h, tape, w, gz, gp, NULL));
    ldeq_tape_free(h, tape, NULL);
    CHECK(ret[5] == LDEQ_RET_MAXITERS && ret[6] == LDEQ_RET_SUCCESS, "retcodes %d %d", ret[5], ret[6]);
    for (int k = 0; k < T; ++k) CHECK(isnan(traj2[(k * B + 5) * 2]) && !isnan(traj2[(k * B + 6) * 2]), "NaN block rule at k = %d", k);
    CHECK(gz[10] == 0.0f && gz[11] == 0.0f && gp[5] == 0.0f && gz[12] != 0.0f, "failed trajectory must have a zero gradient");

    /* usage errors come back as codes with a message, never as a crash */
    CHECK(ldeq_solve_fwd(h, rhs, LDEQ_F32, NULL, d_th, t, B, T, &o, d_traj, NULL, NULL, NULL, NULL, NULL) == LDEQ_ERR_INVALID, "null z0");
    CHECK(strlen(ldeq_last_error(h)) > 0, "no error text");
    o.solver = 7;
    CHECK(ldeq_solve_fwd(h, rhs, LDEQ_F32, d_z0, d_th, t, B, T, &o, d_traj, NULL, NULL, NULL, NULL, NULL) == LDEQ_ERR_UNSUPPORTED, "solver");

    CHECK(ldeq_launch_count(h) > 0, "no kernel launches counted");
    ldeq_rhs_free(h, rhs);
    cudaFree(d_z0); cudaFree(d_th); cudaFree(d_traj); cudaFree(d_w); cudaFree(d_dz0); cudaFree(d_dth); cudaFree(d_ret); cudaFree(d_na);
    ldeq_destroy(h);
    printf("cabi_smoke ok\n");
    return 0;
}
