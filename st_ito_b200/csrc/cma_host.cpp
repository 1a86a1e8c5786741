// Native (host, fp64) CMA-ES behind the C ABI: the sampler / updater that the reference's host loop obtains from pycma
// (st_ito/style_transfer.py:614 `cma.CMAEvolutionStrategy(w0, sigma0, {"bounds": [0, 1], "popsize": P})`, :624 ask, :651
// tell, :639-640 / :672-673 result).  Same algorithm as st_ito_b200/cma.py -- (mu/mu_w, lambda)-CMA-ES with rank-one and
// rank-mu updates, cumulative step-size adaptation and pycma's smooth box transform -- so that a B200 generation
// (13 ms at P = 64, 2-3 ms per GPU when the population is sharded over 8) is not followed by ~0.7 ms of numpy.
// SURVEY 8f-2 ("moving CMA-ES sampling/update off the Python path").  Random numbers: counter-based splitmix64 ->
// Box-Muller (documented in include/stito.h), i.e. a different stream than numpy's; parity of this project is defined on
// evaluate(W) for a given W.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

#include "stito.h"

namespace {

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Box {
    bool on = false;
    double lb = 0, ub = 1, al = 0, au = 0;
    void init(double lo, double hi) {
        on = true; lb = lo; ub = hi;
        const double span = hi - lo;
        al = std::min(span / 2.0, (1.0 + std::fabs(lo)) / 20.0);
        au = std::min(span / 2.0, (1.0 + std::fabs(hi)) / 20.0);
    }
    // pycma BoxConstraintsLinQuadTransformation: identity inside [lb + al, ub - au], quadratic near the bounds,
    // mirrored / periodic outside
    double fwd(double x) const {
        const double span = ub - lb;
        if (x < lb - 2 * al - span / 2.0 || x > ub + 2 * au + span / 2.0) {
            const double r = 2 * (span + al + au), s = lb - 2 * al - span / 2.0;
            x -= r * std::floor((x - s) / r);
        }
        if (x > ub + au) x -= 2 * (x - ub - au);
        if (x < lb - al) x += 2 * (lb - al - x);
        double y = x;
        if (x < lb + al) y = lb + (x - (lb - al)) * (x - (lb - al)) / 4.0 / al;
        else if (x >= ub - au) y = ub - (x - (ub + au)) * (x - (ub + au)) / 4.0 / au;
        return std::min(std::max(y, lb), ub);
    }
    double inv(double y) const {
        y = std::min(std::max(y, lb), ub);
        if (y < lb + al) return (lb - al) + 2.0 * std::sqrt(al * (y - lb));
        if (y > ub - au) return (ub + au) - 2.0 * std::sqrt(au * (ub - y));
        return y;
    }
};

}  // namespace

struct stito_cma {
    int N = 0, lam = 0, mu = 0;
    double sigma = 0, mueff = 0, cc = 0, cs = 0, c1 = 0, cmu = 0, damps = 0, chiN = 0;
    std::vector<double> xmean, weights, pc, ps, B, Dv, C, invsqrtC, geno, best_x, last_f, tmp;
    Box box;
    uint64_t key = 0, draws = 0;
    int64_t countevals = 0, countiter = 0, eigen_at = 0, best_evals = 0;
    double best_f = std::numeric_limits<double>::infinity();
    bool asked = false, has_best = false;

    // The Gaussian draws of the NEXT ask do not depend on tell(): a worker thread produces them while the caller is busy
    // evaluating the population (928 log / sqrt / sincos evaluations ~ 0.1 ms off the critical path).
    std::vector<double> znext;
    std::thread worker;
    std::mutex mu_z;
    std::condition_variable cv_z;
    enum { kIdle, kRequested, kReady, kQuit } zstate = kIdle;

    void worker_loop() {
        std::unique_lock<std::mutex> lock(mu_z);
        for (;;) {
            cv_z.wait(lock, [&] { return zstate == kRequested || zstate == kQuit; });
            if (zstate == kQuit) return;
            lock.unlock();
            normals(znext.data(), znext.size());
            lock.lock();
            zstate = kReady;
            cv_z.notify_all();
        }
    }
    void request_draws() {
        { std::lock_guard<std::mutex> lock(mu_z); zstate = kRequested; }
        cv_z.notify_all();
    }
    void take_draws(std::vector<double> &z) {  // z <- the pre-drawn block (drawing it here if nobody did)
        std::unique_lock<std::mutex> lock(mu_z);
        if (zstate == kIdle) { lock.unlock(); normals(znext.data(), znext.size()); lock.lock(); }
        else cv_z.wait(lock, [&] { return zstate == kReady; });
        z.swap(znext);
        zstate = kIdle;
    }
    ~stito_cma() {
        if (worker.joinable()) {
            { std::unique_lock<std::mutex> lock(mu_z); cv_z.wait(lock, [&] { return zstate != kRequested; }); zstate = kQuit; }
            cv_z.notify_all();
            worker.join();
        }
    }

    // two standard normals per call pair: Box-Muller on uniforms hashed from (key, counter)
    void normals(double *out, size_t n) {
        for (size_t i = 0; i < n; i += 2) {
            const uint64_t a = splitmix64(key + draws), b = splitmix64(key + draws + 1);
            draws += 2;
            const double u1 = ((double)(a >> 11) + 1.0) * 0x1.0p-53, u2 = (double)(b >> 11) * 0x1.0p-53;
            const double r = std::sqrt(-2.0 * std::log(u1)), t = 2.0 * M_PI * u2;
            out[i] = r * std::cos(t);
            if (i + 1 < n) out[i + 1] = r * std::sin(t);
        }
    }

    // C = B diag(D^2) B^T: Householder tridiagonalisation + implicit QL (the EISPACK tred2 / tql2 pair, the same routines
    // pycma carries in pure Python), ~6 N^3 flops; eigenvalues unsorted (the sampler does not care).
    void eigen() {
        sym_eig(C.data(), N, B.data(), Dv.data());
        for (int i = 0; i < N; ++i) Dv[i] = std::sqrt(std::max(Dv[i], 1e-30));
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                double v = 0.0;
                for (int k = 0; k < N; ++k) v += B[(size_t)i * N + k] / Dv[k] * B[(size_t)j * N + k];
                invsqrtC[(size_t)i * N + j] = v;
            }
    }

    static void sym_eig(const double *Ain, int n, double *V, double *d);
};

// V (row-major n x n) <- eigenvectors in columns, d <- eigenvalues of the symmetric matrix A
void stito_cma::sym_eig(const double *Ain, int n, double *V, double *d) {
    std::vector<double> ev(n, 0.0);
    double *e = ev.data();
#define VV(i, j) V[(size_t)(i) * n + (j)]
    for (int i = 0; i < n * n; ++i) V[i] = Ain[i];
    // ---- tred2
    for (int j = 0; j < n; ++j) d[j] = VV(n - 1, j);
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h = h - f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) {
                f = d[j];
                VV(j, i) = f;
                g = e[j] + VV(j, j) * f;
                for (int k = j + 1; k <= i - 1; ++k) { g += VV(k, j) * d[k]; e[k] += VV(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; ++k) VV(k, j) -= (f * e[k] + g * d[k]);
                d[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; ++i) {
        VV(n - 1, i) = VV(i, i);
        VV(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = VV(k, i + 1) / h;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                for (int k = 0; k <= i; ++k) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; ++k) VV(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) VV(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    VV(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
    // ---- tql2
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 0x1.0p-52;
    for (int l = 0; l < n; ++l) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; ++m; }
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 200) break;  // never observed; guards against a NaN input spinning forever
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; ++k) {
                        h = VV(k, i + 1);
                        VV(k, i + 1) = s * VV(k, i) + c * h;
                        VV(k, i) = c * VV(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
#undef VV
}

extern "C" {

int stito_cma_create(const double *x0, int D, double sigma0, int popsize, double lower, double upper, uint64_t seed,
                     stito_cma **out) {
    if (!x0 || !out || D <= 0 || !(sigma0 > 0) || popsize < 2) return STITO_EINVAL;
    stito_cma *es = new stito_cma();
    const int N = D;
    es->N = N; es->lam = popsize; es->mu = popsize / 2; es->sigma = sigma0;
    es->xmean.assign(x0, x0 + N);
    if (lower < upper) {
        es->box.init(lower, upper);
        for (double &v : es->xmean) v = es->box.inv(v);
    }
    es->weights.resize(es->mu);
    double wsum = 0.0;
    for (int i = 0; i < es->mu; ++i) { es->weights[i] = std::log(es->mu + 0.5) - std::log((double)(i + 1)); wsum += es->weights[i]; }
    double w2 = 0.0;
    for (double &w : es->weights) { w /= wsum; w2 += w * w; }
    const double me = es->mueff = 1.0 / w2;
    es->cc = (4 + me / N) / (N + 4 + 2 * me / N);
    es->cs = (me + 2) / (N + me + 5);
    es->c1 = 2 / ((N + 1.3) * (N + 1.3) + me);
    es->cmu = std::min(1 - es->c1, 2 * (me - 2 + 1 / me) / ((N + 2.0) * (N + 2.0) + me));
    es->damps = 1 + 2 * std::max(0.0, std::sqrt((me - 1) / (N + 1)) - 1) + es->cs;
    es->chiN = std::sqrt((double)N) * (1 - 1.0 / (4 * N) + 1.0 / (21.0 * N * N));
    es->pc.assign(N, 0.0); es->ps.assign(N, 0.0); es->Dv.assign(N, 1.0);
    es->B.assign((size_t)N * N, 0.0); es->C.assign((size_t)N * N, 0.0); es->invsqrtC.assign((size_t)N * N, 0.0);
    for (int i = 0; i < N; ++i) es->B[(size_t)i * N + i] = es->C[(size_t)i * N + i] = es->invsqrtC[(size_t)i * N + i] = 1.0;
    es->geno.assign((size_t)popsize * N, 0.0);
    es->best_x.assign(N, 0.0);
    es->key = splitmix64(seed);
    es->znext.assign((size_t)popsize * N, 0.0);
    es->worker = std::thread([es] { es->worker_loop(); });
    es->request_draws();
    *out = es;
    return STITO_OK;
}

void stito_cma_destroy(stito_cma *es) { delete es; }

/* The symmetric eigendecomposition the update uses (exported for the tests): A [n][n] -> V (eigenvectors in columns), d */
int stito_cma_eig(const double *A, int n, double *V, double *d) {
    if (!A || !V || !d || n <= 0) return STITO_EINVAL;
    stito_cma::sym_eig(A, n, V, d);
    return STITO_OK;
}

/* X [popsize][D]: the candidates of this generation, inside the box */
int stito_cma_ask(stito_cma *es, double *X) {
    if (!es || !X) return STITO_EINVAL;
    const int N = es->N, lam = es->lam;
    std::vector<double> z((size_t)lam * N);
    es->take_draws(z);
    for (int k = 0; k < lam; ++k) {
        double *g = es->geno.data() + (size_t)k * N;
        const double *zk = z.data() + (size_t)k * N;
        for (int i = 0; i < N; ++i) {  // y = B (D .* z)
            double y = 0.0;
            for (int j = 0; j < N; ++j) y += es->B[(size_t)i * N + j] * (es->Dv[j] * zk[j]);
            g[i] = es->xmean[i] + es->sigma * y;
            X[(size_t)k * N + i] = es->box.on ? es->box.fwd(g[i]) : g[i];
        }
    }
    es->znext.swap(z);  // hand the buffer back and have the next generation's draws produced in the background
    es->request_draws();
    es->asked = true;
    return STITO_OK;
}

/* genotypes (pre-box-transform search points) of the last ask, [popsize][D] (tests: the update is a function of these) */
int stito_cma_geno(const stito_cma *es, double *G) {
    if (!es || !G) return STITO_EINVAL;
    std::memcpy(G, es->geno.data(), es->geno.size() * sizeof(double));
    return STITO_OK;
}

/* X: the solutions handed out by the preceding ask (only the best one is kept), f [popsize] */
int stito_cma_tell(stito_cma *es, const double *X, const double *fin) {
    if (!es || !X || !fin) return STITO_EINVAL;
    if (!es->asked) return STITO_ESTATE;
    const int N = es->N, lam = es->lam, mu = es->mu;
    std::vector<double> f(fin, fin + lam);
    for (double &v : f) if (!std::isfinite(v)) v = std::numeric_limits<double>::infinity();
    es->countevals += lam;
    es->countiter += 1;
    std::vector<int> order(lam);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return f[a] < f[b]; });
    if (f[order[0]] < es->best_f) {
        es->best_f = f[order[0]];
        std::memcpy(es->best_x.data(), X + (size_t)order[0] * N, N * sizeof(double));
        es->best_evals = es->countevals - lam + order[0] + 1;
        es->has_best = true;
    }
    es->last_f.resize(lam);
    for (int i = 0; i < lam; ++i) es->last_f[i] = f[order[i]];
    std::vector<double> xold = es->xmean, ymean(N), artmp((size_t)mu * N);
    for (int i = 0; i < N; ++i) {
        double m = 0.0;
        for (int k = 0; k < mu; ++k) m += es->weights[k] * es->geno[(size_t)order[k] * N + i];
        es->xmean[i] = m;
        ymean[i] = (m - xold[i]) / es->sigma;
    }
    const double csn = std::sqrt(es->cs * (2 - es->cs) * es->mueff);
    double psn = 0.0;
    for (int i = 0; i < N; ++i) {
        double v = 0.0;
        for (int j = 0; j < N; ++j) v += es->invsqrtC[(size_t)i * N + j] * ymean[j];
        es->ps[i] = (1 - es->cs) * es->ps[i] + csn * v;
        psn += es->ps[i] * es->ps[i];
    }
    psn = std::sqrt(psn);
    const double hsig = (psn / std::sqrt(1 - std::pow(1 - es->cs, 2.0 * es->countiter)) / es->chiN < 1.4 + 2.0 / (N + 1)) ? 1.0 : 0.0;
    const double ccn = std::sqrt(es->cc * (2 - es->cc) * es->mueff);
    for (int i = 0; i < N; ++i) es->pc[i] = (1 - es->cc) * es->pc[i] + hsig * ccn * ymean[i];
    for (int k = 0; k < mu; ++k)
        for (int i = 0; i < N; ++i) artmp[(size_t)k * N + i] = (es->geno[(size_t)order[k] * N + i] - xold[i]) / es->sigma;
    const double keep = 1 - es->c1 - es->cmu, c1a = es->c1 * (1 - hsig) * es->cc * (2 - es->cc);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double rmu = 0.0;
            for (int k = 0; k < mu; ++k) rmu += es->weights[k] * artmp[(size_t)k * N + i] * artmp[(size_t)k * N + j];
            double &c = es->C[(size_t)i * N + j];
            c = keep * c + es->c1 * es->pc[i] * es->pc[j] + c1a * c + es->cmu * rmu;
        }
    es->sigma *= std::exp((es->cs / es->damps) * (psn / es->chiN - 1));
    if ((double)(es->countevals - es->eigen_at) > lam / (es->c1 + es->cmu) / N / 10) {
        es->eigen_at = es->countevals;
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) es->C[(size_t)j * N + i] = es->C[(size_t)i * N + j];
        es->eigen();
    }
    es->asked = false;
    return STITO_OK;
}

/* best-ever solution / value (has_best = 0 before the first tell), distribution mean mapped into the box, sigma,
 * per-coordinate standard deviations, counters.  Every output pointer is nullable. */
int stito_cma_result(const stito_cma *es, double *xbest, double *fbest, int *has_best, double *xfavorite, double *sigma,
                     double *stds, int64_t *evals_best, int64_t *evaluations, int64_t *iterations, double *axis_ratio,
                     double *last_f_range) {
    if (!es) return STITO_EINVAL;
    const int N = es->N;
    if (xbest) std::memcpy(xbest, es->best_x.data(), N * sizeof(double));
    if (fbest) *fbest = es->best_f;
    if (has_best) *has_best = es->has_best ? 1 : 0;
    if (xfavorite) for (int i = 0; i < N; ++i) xfavorite[i] = es->box.on ? es->box.fwd(es->xmean[i]) : es->xmean[i];
    if (sigma) *sigma = es->sigma;
    if (stds) for (int i = 0; i < N; ++i) stds[i] = es->sigma * std::sqrt(std::max(es->C[(size_t)i * N + i], 0.0));
    if (evals_best) *evals_best = es->best_evals;
    if (evaluations) *evaluations = es->countevals;
    if (iterations) *iterations = es->countiter;
    if (axis_ratio) *axis_ratio = *std::max_element(es->Dv.begin(), es->Dv.end()) / *std::min_element(es->Dv.begin(), es->Dv.end());
    if (last_f_range) {
        *last_f_range = es->last_f.empty() ? -1.0
                                           : *std::max_element(es->last_f.begin(), es->last_f.end()) -
                                                 *std::min_element(es->last_f.begin(), es->last_f.end());
    }
    return STITO_OK;
}

}  // extern "C"
