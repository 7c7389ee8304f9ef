/*
 * C restatement of the OQRL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.  The
 * product (oqrl_b200/) never links or calls it.
 *
 * Follows, gate by gate in float64:
 *   /root/reference/src/agent.py:36-46   circuit (AngleEmbedding RY + StronglyEntanglingLayers + <Z_i>)
 *   /root/reference/src/agent.py:48-61   one circuit per batch row
 *   /root/reference/src/optim.py:152-173 TD target with the candidate's own weights, gamma 0.99, -MSE
 *   /root/reference/src/optim.py:204     one evaluation per candidate (the OpenMP loop here)
 * with the PennyLane template semantics of SURVEY.md Appendix A (PennyLane
 * itself -- requirements.txt:7, unpinned -- is not installable here:
 * PARITY UNPINNED by the reference; pinned instead against
 * tests/golden/kat_appendix_b.json and against oracle/pqc_oracle.py).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t n_qubits, n_layers, n_out, n_feat;
    int32_t entangler;   /* 0 = SEL_CNOT (reference), 1 = CZ_RING */
    int32_t reupload;    /* 0/1 */
    int32_t feature_map; /* 0 = first_k (AngleEmbedding), 1 = tiled */
    int32_t reserved;
} oracle_spec;

typedef struct { double re, im; } cplx;

static inline int bit_of(uint64_t k, int wire, int n) { return (int)((k >> (n - 1 - wire)) & 1u); }

/* general 2x2 complex gate on `wire` */
static void apply_1q(cplx *psi, int n, int wire, const cplx m[4])
{
    const uint64_t dim = 1ull << n, stride = 1ull << (n - 1 - wire);
    for (uint64_t base = 0; base < dim; base += 2 * stride)
        for (uint64_t j = base; j < base + stride; ++j) {
            cplx a = psi[j], b = psi[j + stride];
            psi[j].re = m[0].re * a.re - m[0].im * a.im + m[1].re * b.re - m[1].im * b.im;
            psi[j].im = m[0].re * a.im + m[0].im * a.re + m[1].re * b.im + m[1].im * b.re;
            psi[j + stride].re = m[2].re * a.re - m[2].im * a.im + m[3].re * b.re - m[3].im * b.im;
            psi[j + stride].im = m[2].re * a.im + m[2].im * a.re + m[3].re * b.im + m[3].im * b.re;
        }
}

static void apply_cnot(cplx *psi, int n, int control, int target)
{
    const uint64_t dim = 1ull << n, tbit = 1ull << (n - 1 - target);
    for (uint64_t k = 0; k < dim; ++k)
        if (bit_of(k, control, n) && !(k & tbit)) {
            cplx t = psi[k]; psi[k] = psi[k | tbit]; psi[k | tbit] = t;
        }
}

static void apply_cz(cplx *psi, int n, int a, int b)
{
    const uint64_t dim = 1ull << n;
    for (uint64_t k = 0; k < dim; ++k)
        if (bit_of(k, a, n) && bit_of(k, b, n)) { psi[k].re = -psi[k].re; psi[k].im = -psi[k].im; }
}

static void ry_gate(double t, cplx m[4])
{
    double c = cos(0.5 * t), s = sin(0.5 * t);
    m[0].re = c; m[0].im = 0; m[1].re = -s; m[1].im = 0;
    m[2].re = s; m[2].im = 0; m[3].re = c;  m[3].im = 0;
}

/* Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi) */
static void rot_gate(double phi, double theta, double omega, cplx m[4])
{
    double c = cos(0.5 * theta), s = sin(0.5 * theta);
    double p = 0.5 * (phi + omega), q = 0.5 * (phi - omega);
    m[0].re = cos(p) * c;  m[0].im = -sin(p) * c;
    m[1].re = -cos(q) * s; m[1].im = -sin(q) * s;
    m[2].re = cos(q) * s;  m[2].im = -sin(q) * s;
    m[3].re = cos(p) * c;  m[3].im = sin(p) * c;
}

static void embed(cplx *psi, const oracle_spec *sp, const float *x)
{
    cplx m[4];
    const int n = sp->n_qubits;
    if (sp->n_feat <= 0) return;
    if (sp->feature_map == 1) {
        for (int w = 0; w < n; ++w) { ry_gate((double)x[w % sp->n_feat], m); apply_1q(psi, n, w, m); }
    } else {
        for (int w = 0; w < sp->n_feat && w < n; ++w) { ry_gate((double)x[w], m); apply_1q(psi, n, w, m); }
    }
}

/* one circuit: params[D] f32, x[n_feat] f32 -> q[n_out] f64; psi is scratch of 2^n */
static void circuit(const oracle_spec *sp, const float *params, const float *x, cplx *psi, double *q)
{
    const int n = sp->n_qubits, L = sp->n_layers;
    const uint64_t dim = 1ull << n;
    cplx m[4];
    memset(psi, 0, dim * sizeof(cplx));
    psi[0].re = 1.0;
    if (!sp->reupload || L == 0) embed(psi, sp, x);
    for (int l = 0; l < L; ++l) {
        if (sp->reupload) embed(psi, sp, x);
        for (int i = 0; i < n; ++i) {
            const float *w = params + ((size_t)l * n + i) * 3;
            rot_gate((double)w[0], (double)w[1], (double)w[2], m);
            apply_1q(psi, n, i, m);
        }
        if (n > 1) {
            if (sp->entangler == 0) {
                int r = (l % (n - 1)) + 1;
                for (int i = 0; i < n; ++i) apply_cnot(psi, n, i, (i + r) % n);
            } else {
                for (int i = 0; i < n; ++i) apply_cz(psi, n, i, (i + 1) % n);
            }
        }
    }
    for (int i = 0; i < sp->n_out; ++i) q[i] = 0.0;
    for (uint64_t k = 0; k < dim; ++k) {
        double p = psi[k].re * psi[k].re + psi[k].im * psi[k].im;
        for (int i = 0; i < sp->n_out; ++i) q[i] += bit_of(k, i, n) ? -p : p;
    }
}

static int check_spec(const oracle_spec *sp)
{
    if (!sp || sp->n_qubits < 1 || sp->n_qubits > 26 || sp->n_layers < 0) return -1;
    if (sp->n_out < 1 || sp->n_out > sp->n_qubits || sp->n_out > 32) return -1;
    if (sp->feature_map == 0 && sp->n_feat > sp->n_qubits) return -1;
    return 0;
}

/* q[P][B][n_out] (f64) for P parameter vectors on B input rows */
int oracle_qvalues(const oracle_spec *sp, const float *params, int P, const float *x, int B,
                   double *q, int nthreads)
{
    if (check_spec(sp)) return -1;
    const int D = sp->n_layers * sp->n_qubits * 3;
    const size_t dim = (size_t)1 << sp->n_qubits;
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cplx *psi = (cplx *)malloc(dim * sizeof(cplx));
        if (!psi) { err = -2; }
        else {
#pragma omp for schedule(static) collapse(2)
            for (int p = 0; p < P; ++p)
                for (int b = 0; b < B; ++b)
                    circuit(sp, params + (size_t)p * D, x + (size_t)b * sp->n_feat, psi,
                            q + ((size_t)p * B + b) * sp->n_out);
            free(psi);
        }
    }
    return err;
}

/* fitness[P] (f64) = -mean_t (Q(s_t)[a_t] - (r_t + (1-d_t)*gamma*max_a Q(s'_t)[a]))^2 */
int oracle_fitness(const oracle_spec *sp, const float *params, int P,
                   const float *obs, const float *next_obs, const int32_t *act,
                   const float *rew, const float *term, int T, double gamma,
                   double *fitness, int nthreads)
{
    if (check_spec(sp) || T <= 0) return -1;
    const int D = sp->n_layers * sp->n_qubits * 3;
    const size_t dim = (size_t)1 << sp->n_qubits;
    for (int t = 0; t < T; ++t) if (act[t] < 0 || act[t] >= sp->n_out) return -3;
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cplx *psi = (cplx *)malloc(dim * sizeof(cplx));
        double q[32], qn[32];
        if (!psi) { err = -2; }
        else {
#pragma omp for schedule(dynamic, 1)
            for (int p = 0; p < P; ++p) {
                double acc = 0.0;
                for (int t = 0; t < T; ++t) {
                    circuit(sp, params + (size_t)p * D, obs + (size_t)t * sp->n_feat, psi, q);
                    circuit(sp, params + (size_t)p * D, next_obs + (size_t)t * sp->n_feat, psi, qn);
                    double mx = qn[0];
                    for (int i = 1; i < sp->n_out; ++i) if (qn[i] > mx) mx = qn[i];
                    /* optim.py:163: (1 - d) * 0.99 is float32 arithmetic in the reference */
                    float keep = (1.0f - term[t]) * (float)gamma;
                    double d = q[act[t]] - ((double)rew[t] + (double)keep * mx);
                    acc += d * d;
                }
                fitness[p] = -acc / (double)T;
            }
            free(psi);
        }
    }
    return err;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
