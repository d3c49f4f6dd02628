// TEST INFRASTRUCTURE ONLY — C entry points of the CPU oracle (ctypes-loaded by oracle/oracle.py).
// Element kinds follow include/gschur_cuda.h: 0 = f64, 1 = c64, 2 = dd (hi,lo), 3 = complex dd.
// Matrices are column-major.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>
#include "gschur_oracle.hpp"

using namespace gso;

namespace {

template <class R> struct Conv;   // storage <-> scalar conversions per kind
template <> struct Conv<double> {
    static const int nd = 1;
    static double load(const double* p) { return p[0]; }
    static void store(double* p, double v) { p[0] = v; }
};
template <> struct Conv<DD> {
    static const int nd = 2;
    static DD load(const double* p) { return DD(p[0], p[1]); }
    static void store(double* p, const DD& v) { p[0] = v.hi; p[1] = v.lo; }
};

// load a kind-typed matrix into MP (exact)
static void load_mp(int kind, int n, const void* A, long lda, std::vector<MP>& re, std::vector<MP>& im) {
    const double* a = (const double*)A;
    bool cplx = (kind & 1);
    int nd = (kind >= 2) ? 2 : 1;
    int stride = nd * (cplx ? 2 : 1);
    re.assign((size_t)n * n, MP(0.0));
    if (cplx) im.assign((size_t)n * n, MP(0.0));
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const double* e = a + ((size_t)i + (size_t)j * lda) * stride;
            if (nd == 1) {
                re[i + (size_t)j * n] = MP(e[0]);
                if (cplx) im[i + (size_t)j * n] = MP(e[1]);
            } else {
                re[i + (size_t)j * n] = MP(DD(e[0], e[1]));
                if (cplx) im[i + (size_t)j * n] = MP(DD(e[2], e[3]));
            }
        }
}

template <class R> int run_real(int n, double* A, long lda, double* Z, long ldz, double* w, int scale, int maxiter,
                                long* stats) {
    Mat<R> Am((R*)A, n, n, lda);
    Mat<R> Zm((R*)Z, n, n, ldz);
    Stats st;
    int rc;
    try {
        rc = gschur_real<R>(Am, Z ? &Zm : nullptr, (Cx<R>*)w, scale != 0, maxiter, &st);
    } catch (Unconverged&) {
        rc = 1;
    }
    if (stats) {
        stats[0] = st.sweeps;
        stats[1] = st.applications;
        stats[2] = st.rowpairs;
        stats[3] = st.exceptional;
    }
    return rc;
}
template <class R> int run_complex(int n, double* A, long lda, double* Z, long ldz, double* w, int scale,
                                   int maxiter, long* stats) {
    Mat<Cx<R>> Am((Cx<R>*)A, n, n, lda);
    Mat<Cx<R>> Zm((Cx<R>*)Z, n, n, ldz);
    Stats st;
    int rc;
    try {
        rc = gschur_complex<R>(Am, Z ? &Zm : nullptr, (Cx<R>*)w, scale != 0, maxiter, &st);
    } catch (Unconverged&) {
        rc = 1;
    }
    if (stats) {
        stats[0] = st.sweeps;
        stats[1] = st.applications;
        stats[2] = st.rowpairs;
        stats[3] = st.exceptional;
    }
    return rc;
}

}  // namespace

extern "C" {

// gschur!(A; wantZ = (Z != NULL), scale) in the arithmetic of `kind`.  A is overwritten by T, w gets n complex
// eigenvalues.  Returns 0, 1 = UnconvergedException, -1 = bad argument.
int gso_gschur(int kind, int n, void* A, long lda, void* Z, long ldz, void* w, int scale, int maxiter, long* stats) {
    if (n < 0 || lda < n || (Z && ldz < n)) return -1;
    static_assert(sizeof(DD) == 16 && sizeof(Cx<double>) == 16 && sizeof(Cx<DD>) == 32, "layout");
    switch (kind) {
        case 0: return run_real<double>(n, (double*)A, lda, (double*)Z, ldz, (double*)w, scale, maxiter, stats);
        case 1: return run_complex<double>(n, (double*)A, lda, (double*)Z, ldz, (double*)w, scale, maxiter, stats);
        case 2: return run_real<DD>(n, (double*)A, lda, (double*)Z, ldz, (double*)w, scale, maxiter, stats);
        case 3: return run_complex<DD>(n, (double*)A, lda, (double*)Z, ldz, (double*)w, scale, maxiter, stats);
    }
    return -1;
}

// Same decomposition carried out in MPFR 256-bit arithmetic (the stand-in for BigFloat(256)).  Input of `kind`
// is converted exactly; outputs are rounded to double-double: T, Z as (2 or 4, n, n) arrays, w as (4, n).
int gso_gschur_mp(int kind, int n, const void* A, long lda, double* Tdd, double* Zdd, double* wdd, int scale) {
    if (n < 0 || lda < n) return -1;
    std::vector<MP> re, im;
    load_mp(kind, n, A, lda, re, im);
    bool cplx = kind & 1;
    int rc = 0;
    try {
        if (!cplx) {
            std::vector<MP> Zs((size_t)n * n);
            std::vector<Cx<MP>> w(n);
            Mat<MP> Am(re.data(), n, n, n), Zm(Zs.data(), n, n, n);
            rc = gschur_real<MP>(Am, &Zm, w.data(), scale != 0, 0, nullptr);
            for (size_t i = 0; i < (size_t)n * n; ++i) {
                DD t = re[i].to_dd(), z = Zs[i].to_dd();
                Tdd[2 * i] = t.hi; Tdd[2 * i + 1] = t.lo;
                Zdd[2 * i] = z.hi; Zdd[2 * i + 1] = z.lo;
            }
            for (int i = 0; i < n; ++i) {
                DD a = w[i].re.to_dd(), b = w[i].im.to_dd();
                wdd[4 * i] = a.hi; wdd[4 * i + 1] = a.lo; wdd[4 * i + 2] = b.hi; wdd[4 * i + 3] = b.lo;
            }
        } else {
            std::vector<Cx<MP>> As((size_t)n * n), Zs((size_t)n * n), w(n);
            for (size_t i = 0; i < (size_t)n * n; ++i) As[i] = Cx<MP>(re[i], im[i]);
            Mat<Cx<MP>> Am(As.data(), n, n, n), Zm(Zs.data(), n, n, n);
            rc = gschur_complex<MP>(Am, &Zm, w.data(), scale != 0, 0, nullptr);
            for (size_t i = 0; i < (size_t)n * n; ++i) {
                DD a = As[i].re.to_dd(), b = As[i].im.to_dd(), c = Zs[i].re.to_dd(), d = Zs[i].im.to_dd();
                Tdd[4 * i] = a.hi; Tdd[4 * i + 1] = a.lo; Tdd[4 * i + 2] = b.hi; Tdd[4 * i + 3] = b.lo;
                Zdd[4 * i] = c.hi; Zdd[4 * i + 1] = c.lo; Zdd[4 * i + 2] = d.hi; Zdd[4 * i + 3] = d.lo;
            }
            for (int i = 0; i < n; ++i) {
                DD a = w[i].re.to_dd(), b = w[i].im.to_dd();
                wdd[4 * i] = a.hi; wdd[4 * i + 1] = a.lo; wdd[4 * i + 2] = b.hi; wdd[4 * i + 3] = b.lo;
            }
        }
    } catch (Unconverged&) {
        rc = 1;
    }
    return rc;
}

// The reference's acceptance ratios (test/complex.jl:15,18; test/real.jl:40,43), evaluated in MPFR 256:
//   out[0] = ||A - Z T Z'||_F / (n * ||A||_F * ulp)      out[1] = ||I - Z Z'||_F / (n * ulp)
// with ulp = eps of `kind` (2^-52 for kinds 0/1, 2^-104 for kinds 2/3).  out[2] = ||A||_F (double).
int gso_residuals(int kind, int n, const void* A, long lda, const void* T, long ldt, const void* Z, long ldz,
                  double* out) {
    if (n <= 0) { out[0] = out[1] = out[2] = 0.0; return 0; }
    bool cplx = kind & 1;
    std::vector<MP> Ar, Ai, Tr, Ti, Zr, Zi;
    load_mp(kind, n, A, lda, Ar, Ai);
    load_mp(kind, n, T, ldt, Tr, Ti);
    load_mp(kind, n, Z, ldz, Zr, Zi);
    size_t nn = (size_t)n * n;
    MP zero(0.0);
    // W = Z * T
    std::vector<MP> Wr(nn, zero), Wi(cplx ? nn : 0, zero);
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k) {
            const MP& tr = Tr[k + (size_t)j * n];
            bool tz = mpfr_zero_p(tr.v);
            if (cplx) {
                const MP& ti = Ti[k + (size_t)j * n];
                if (tz && mpfr_zero_p(ti.v)) continue;
                for (int i = 0; i < n; ++i) {
                    const MP& zr = Zr[i + (size_t)k * n];
                    const MP& zi = Zi[i + (size_t)k * n];
                    Wr[i + (size_t)j * n] += zr * tr - zi * ti;
                    Wi[i + (size_t)j * n] += zr * ti + zi * tr;
                }
            } else {
                if (tz) continue;
                for (int i = 0; i < n; ++i) Wr[i + (size_t)j * n] += Zr[i + (size_t)k * n] * tr;
            }
        }
    // R = A - W * Z'   and   O = I - Z * Z'
    MP res(0.0), orth(0.0), an(0.0);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            MP sr(0.0), si(0.0), orr(0.0), oi(0.0);
            for (int k = 0; k < n; ++k) {
                const MP& zjr = Zr[j + (size_t)k * n];
                if (cplx) {
                    const MP& zji = Zi[j + (size_t)k * n];
                    // W[i,k] * conj(Z[j,k])
                    sr += Wr[i + (size_t)k * n] * zjr + Wi[i + (size_t)k * n] * zji;
                    si += Wi[i + (size_t)k * n] * zjr - Wr[i + (size_t)k * n] * zji;
                    orr += Zr[i + (size_t)k * n] * zjr + Zi[i + (size_t)k * n] * zji;
                    oi += Zi[i + (size_t)k * n] * zjr - Zr[i + (size_t)k * n] * zji;
                } else {
                    sr += Wr[i + (size_t)k * n] * zjr;
                    orr += Zr[i + (size_t)k * n] * zjr;
                }
            }
            MP dr = Ar[i + (size_t)j * n] - sr;
            res += dr * dr;
            an += Ar[i + (size_t)j * n] * Ar[i + (size_t)j * n];
            MP od = (i == j ? MP(1.0) : MP(0.0)) - orr;
            orth += od * od;
            if (cplx) {
                MP di = Ai[i + (size_t)j * n] - si;
                res += di * di;
                an += Ai[i + (size_t)j * n] * Ai[i + (size_t)j * n];
                orth += oi * oi;
            }
        }
    MP ulp = (kind >= 2) ? MP(std::ldexp(1.0, -104)) : MP(std::ldexp(1.0, -52));
    MP nn_((double)n);
    MP anorm = RT<MP>::sqrt(an);
    out[2] = anorm.to_double();
    if (mpfr_zero_p(anorm.v)) out[0] = RT<MP>::sqrt(res).to_double();
    else out[0] = (RT<MP>::sqrt(res) / (nn_ * anorm * ulp)).to_double();
    out[1] = (RT<MP>::sqrt(orth) / (nn_ * ulp)).to_double();
    return 0;
}

// _hessenberg!(A) then _materializeQ: A <- factors (H on and above the sub-diagonal, reflector tails below),
// tau[n-1], Q (nullable) n x n.
int gso_hessenberg(int kind, int n, void* A, long lda, void* tau, void* Q, long ldq) {
    if (n < 0 || lda < n || (Q && ldq < n)) return -1;
    if (n == 0) return 0;
    switch (kind) {
        case 0: {
            Mat<double> Am((double*)A, n, n, lda), Qm((double*)Q, n, n, ldq);
            hessenberg(Am, (double*)tau);
            if (Q) materializeQ(Am, (const double*)tau, Qm);
            return 0;
        }
        case 1: {
            Mat<Cx<double>> Am((Cx<double>*)A, n, n, lda), Qm((Cx<double>*)Q, n, n, ldq);
            hessenberg(Am, (Cx<double>*)tau);
            if (Q) materializeQ(Am, (const Cx<double>*)tau, Qm);
            return 0;
        }
        case 2: {
            Mat<DD> Am((DD*)A, n, n, lda), Qm((DD*)Q, n, n, ldq);
            hessenberg(Am, (DD*)tau);
            if (Q) materializeQ(Am, (const DD*)tau, Qm);
            return 0;
        }
        case 3: {
            Mat<Cx<DD>> Am((Cx<DD>*)A, n, n, lda), Qm((Cx<DD>*)Q, n, n, ldq);
            hessenberg(Am, (Cx<DD>*)tau);
            if (Q) materializeQ(Am, (const Cx<DD>*)tau, Qm);
            return 0;
        }
    }
    return -1;
}

// gschur!(H::Hessenberg, Z) entry: H upper Hessenberg in place.  Returns -2 for a non-real sub-diagonal
// (ArgumentError, src/GenericSchur.jl:206-210), 1 for UnconvergedException.
int gso_gschur_hess(int kind, int n, void* H, long ldh, void* Z, long ldz, void* w, int maxiter, int checksd) {
    if (n < 0 || ldh < n || (Z && ldz < n)) return -1;
    try {
        if (kind == 0) {
            Mat<double> Hm((double*)H, n, n, ldh), Zm((double*)Z, n, n, ldz);
            return gschur_hess_real<double>(Hm, Z ? &Zm : nullptr, (Cx<double>*)w, maxiter, nullptr);
        } else if (kind == 1) {
            Mat<Cx<double>> Hm((Cx<double>*)H, n, n, ldh), Zm((Cx<double>*)Z, n, n, ldz);
            int rc = gschur_hess_complex<double>(Hm, Z ? &Zm : nullptr, (Cx<double>*)w, maxiter, 0, checksd != 0, nullptr);
            return rc == -1 ? -2 : rc;
        }
    } catch (Unconverged&) {
        return 1;
    }
    return -1;
}

// reciprocal eigenvalue condition numbers of an upper triangular complex T (kind 1 or 3), as doubles
int gso_eigvalscond(int kind, int n, const void* T, long ldt, double* s) {
    if (kind == 1) {
        Mat<Cx<double>> Tm((Cx<double>*)T, n, n, ldt);
        eigvalscond_triu<double>(Tm, s);
        return 0;
    } else if (kind == 3) {
        Mat<Cx<DD>> Tm((Cx<DD>*)T, n, n, ldt);
        std::vector<DD> sd(n);
        eigvalscond_triu<DD>(Tm, sd.data());
        for (int i = 0; i < n; ++i) s[i] = sd[i].hi;
        return 0;
    }
    return -1;
}

// _gs2x2!: abcd = (a, b, c, d) in place, csn = (cs, sn), w = (re1, im1, re2, im2)
int gso_gs2x2(double* abcd, double* csn, double* w) {
    Cx<double> w1, w2;
    gs2x2<double>(abcd[0], abcd[1], abcd[2], abcd[3], csn[0], csn[1], w1, w2);
    w[0] = w1.re; w[1] = w1.im; w[2] = w2.re; w[3] = w2.im;
    return 0;
}

// _reflector!(x): x[0..n-1] in place, tau out (1 or 2 doubles); kinds 0 and 1
int gso_reflector(int kind, int n, void* x, void* tau) {
    if (kind == 0) { *(double*)tau = reflector<double>((double*)x, n); return 0; }
    if (kind == 1) { *(Cx<double>*)tau = reflector<double>((Cx<double>*)x, n); return 0; }
    return -1;
}

// gschur! over a contiguous batch (strides n*n and n) with `nthreads` host threads: the CPU baseline of bench.py.
int gso_gschur_batched(int kind, int n, long batch, void* A, void* Z, void* w, int scale, int nthreads, int* info) {
    int esz = (kind == 0) ? 8 : (kind == 3 ? 32 : 16);
    int wsz = (kind >= 2) ? 32 : 16;
    if (nthreads < 1) nthreads = 1;
    std::atomic<long> next(0);
    std::atomic<int> bad(0);
    auto worker = [&]() {
        for (;;) {
            long b = next.fetch_add(1);
            if (b >= batch) break;
            char* Ab = (char*)A + (size_t)b * n * n * esz;
            char* Zb = Z ? (char*)Z + (size_t)b * n * n * esz : nullptr;
            char* wb = (char*)w + (size_t)b * n * wsz;
            int rc = gso_gschur(kind, n, Ab, n, Zb, n, wb, scale, 0, nullptr);
            if (info) info[b] = rc;
            if (rc) bad.fetch_add(1);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return bad.load();
}

// balance!(A; scale, permute): A in place; D (n doubles), sp (n ints), ii = (ilo, ihi, trivial); kinds 0 and 1
int gso_balance(int kind, int n, void* A, long lda, int scale, int permute, double* D, int* sp, int* ii) {
    int ilo = 1, ihi = n, rc = -1;
    bool trivial = true;
    if (kind == 0) rc = balance<double>(Mat<double>((double*)A, n, n, lda), scale != 0, permute != 0, D, sp, ilo, ihi, trivial);
    else if (kind == 1) rc = balance<Cx<double>>(Mat<Cx<double>>((Cx<double>*)A, n, n, lda), scale != 0, permute != 0, D, sp, ilo, ihi, trivial);
    ii[0] = ilo; ii[1] = ihi; ii[2] = trivial ? 1 : 0;
    return rc;
}
// lmul!(B, V) / ldiv!(B, V)
int gso_balance_apply(int kind, int n, void* V, long ldv, const double* D, const int* sp, const int* ii, int inv) {
    if (kind == 0) balance_apply<double>(Mat<double>((double*)V, n, n, ldv), D, sp, ii[0], ii[1], ii[2] != 0, inv != 0);
    else if (kind == 1) balance_apply<Cx<double>>(Mat<Cx<double>>((Cx<double>*)V, n, n, ldv), D, sp, ii[0], ii[1], ii[2] != 0, inv != 0);
    else return -1;
    return 0;
}

int gso_version(void) { return 1; }

}  // extern "C"
