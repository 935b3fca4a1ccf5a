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

    /* the recurrent pattern extractor (GOKU.jl:30-49) and the loss with the sigmoid output layer folded in, through the header
       alone: final states, reverse pass checked by a finite difference on one frame entry; loss(logits) == loss(sigmoid) */
    {
        enum { PB = 40, PT = 9, PF = 32, PH = 16 };
        const int nr = ldeq_pattern_extractor_param_count(0, PF, PH), nl = ldeq_pattern_extractor_param_count(1, PF, PH);
        CHECK(nr == 1344 && nl == 5312, "pattern extractor parameter counts %d %d", nr, nl);
        float *hx = (float*)malloc(sizeof(float) * PT * PB * PF), *hp = (float*)malloc(sizeof(float) * (nr + 2 * nl));
        for (int i = 0; i < PT * PB * PF; ++i) { s = s * 1664525u + 1013904223u; hx[i] = ((s >> 8) / 16777216.0f - 0.5f) * 2.0f; }
        for (int i = 0; i < nr + 2 * nl; ++i) { s = s * 1664525u + 1013904223u; hp[i] = ((s >> 8) / 16777216.0f - 0.5f) * 0.5f; }
        float *dx_, *dp_, *dz_, *dt_, *dgx, *dgp, *dcz, *dct;
        CHECK(cudaMalloc((void**)&dx_, sizeof(float) * PT * PB * PF) == cudaSuccess && cudaMalloc((void**)&dp_, sizeof(float) * (nr + 2 * nl)) == cudaSuccess &&
              cudaMalloc((void**)&dz_, sizeof(float) * PB * PH) == cudaSuccess && cudaMalloc((void**)&dt_, sizeof(float) * PB * 2 * PH) == cudaSuccess &&
              cudaMalloc((void**)&dgx, sizeof(float) * PT * PB * PF) == cudaSuccess && cudaMalloc((void**)&dgp, sizeof(float) * (nr + 2 * nl)) == cudaSuccess &&
              cudaMalloc((void**)&dcz, sizeof(float) * PB * PH) == cudaSuccess && cudaMalloc((void**)&dct, sizeof(float) * PB * 2 * PH) == cudaSuccess, "cudaMalloc");
        cudaMemcpy(dp_, hp, sizeof(float) * (nr + 2 * nl), cudaMemcpyHostToDevice);
        static float hz[PB * PH], ht[PB * 2 * PH], ones_z[PB * PH], ones_t[PB * 2 * PH], hgx[PT * PB * PF];
        for (int i = 0; i < PB * PH; ++i) ones_z[i] = 1.0f;
        for (int i = 0; i < PB * 2 * PH; ++i) ones_t[i] = 1.0f;
        cudaMemcpy(dcz, ones_z, sizeof ones_z, cudaMemcpyHostToDevice);
        cudaMemcpy(dct, ones_t, sizeof ones_t, cudaMemcpyHostToDevice);
        double L[3];
        const int probe = (4 * PB + 7) * PF + 11;   /* frame 4, sequence 7, feature 11 */
        for (int pass = 0; pass < 3; ++pass) {       /* 0: base (+ reverse pass), 1: +eps, 2: -eps */
            const float keep = hx[probe];
            hx[probe] = keep + (pass == 1 ? 1e-2f : pass == 2 ? -1e-2f : 0.0f);
            cudaMemcpy(dx_, hx, sizeof(float) * PT * PB * PF, cudaMemcpyHostToDevice);
            hx[probe] = keep;
            ldeq_pe_tape* pt = NULL;
            LD(ldeq_pattern_extractor_fwd(h, dx_, PB, PT, PF, PH, dp_, dp_ + nr, dp_ + nr + nl, dz_, dt_, pass == 0 ? &pt : NULL, NULL));
            if (pass == 0) {
                CHECK(pt != NULL, "no pattern-extractor tape");
                LD(ldeq_pattern_extractor_bwd(h, pt, dx_, dp_, dp_ + nr, dp_ + nr + nl, dcz, dct, dgx, dgp, dgp + nr, dgp + nr + nl, NULL));
                ldeq_pe_tape_free(h, pt, NULL);
                cudaMemcpy(hgx, dgx, sizeof hgx, cudaMemcpyDeviceToHost);
            }
            cudaMemcpy(hz, dz_, sizeof hz, cudaMemcpyDeviceToHost);
            cudaMemcpy(ht, dt_, sizeof ht, cudaMemcpyDeviceToHost);
            L[pass] = 0.0;
            for (int i = 0; i < PB * PH; ++i) L[pass] += hz[i];
            for (int i = 0; i < PB * 2 * PH; ++i) L[pass] += ht[i];
        }
        const double fd = (L[1] - L[2]) / 2e-2;
        CHECK(isfinite(L[0]) && fabs(fd - hgx[probe]) <= 2e-2 * fmax(fabs(fd), 1e-3), "pattern extractor: finite difference %g vs dx %g", fd, hgx[probe]);
        CHECK(ldeq_pattern_extractor_fwd(h, dx_, PB, PT, 24, PH, dp_, NULL, NULL, dz_, NULL, NULL, NULL) == LDEQ_ERR_UNSUPPORTED, "F = 24 must be refused");
        /* ELBO on pre-activations: reuse dx_ (PT*PB*PF values) as logits a and dgx as the data x in [0,1) */
        const int EP = PF, EB = PB, ET = PT;
        float* hxx = (float*)malloc(sizeof(float) * ET * EB * EP);
        float* hsig = (float*)malloc(sizeof(float) * ET * EB * EP);
        for (int i = 0; i < ET * EB * EP; ++i) { s = s * 1664525u + 1013904223u; hxx[i] = (s >> 8) / 16777216.0f; hsig[i] = 1.0f / (1.0f + expf(-hx[i])); }
        cudaMemcpy(dx_, hx, sizeof(float) * ET * EB * EP, cudaMemcpyHostToDevice);
        cudaMemcpy(dgx, hxx, sizeof(float) * ET * EB * EP, cudaMemcpyHostToDevice);
        float *dl, *dsig;
        CHECK(cudaMalloc((void**)&dl, 6 * sizeof(float)) == cudaSuccess && cudaMalloc((void**)&dsig, sizeof(float) * ET * EB * EP) == cudaSuccess, "cudaMalloc");
        cudaMemcpy(dsig, hsig, sizeof(float) * ET * EB * EP, cudaMemcpyHostToDevice);
        LD(ldeq_elbo_logits_fwd_bwd(h, dgx, dx_, NULL, NULL, NULL, 0, 0.5f, EB, ET, EP, 1.0f, dl, NULL, NULL, NULL, NULL));
        LD(ldeq_elbo_fwd_bwd(h, dgx, dsig, NULL, NULL, NULL, 0, 0.5f, EB, ET, EP, 1.0f, dl + 3, NULL, NULL, NULL, NULL));
        float hl[6];
        cudaMemcpy(hl, dl, sizeof hl, cudaMemcpyDeviceToHost);
        CHECK(hl[0] > 0.0f && fabsf(hl[0] - hl[3]) <= 1e-5f * hl[3], "elbo on logits %g vs elbo on sigmoid(logits) %g", hl[0], hl[3]);
        free(hx); free(hp); free(hxx); free(hsig);
        cudaFree(dx_); cudaFree(dp_); cudaFree(dz_); cudaFree(dt_); cudaFree(dgx); cudaFree(dgp); cudaFree(dcz); cudaFree(dct); cudaFree(dl); cudaFree(dsig);
    }

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
