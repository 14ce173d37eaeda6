/* pg_perlnum.h -- numeric value of a string the way Perl's numeric operators see it
 * (Megaclust/megaclust2.pl:126-130 compares text fields with `<` and `>`): optional blanks,
 * optional sign, digits with an optional decimal point, optional exponent, trailing garbage
 * ignored, "inf"/"infinity"/"nan" in any case, everything else 0.  Shared by the CUDA kernels
 * and the C command-line tools so that a threshold given on the command line and a field of
 * the input parse to the same double.
 *
 * The mantissa is taken to 19 significant digits (more are dropped with their weight kept),
 * then scaled by a power of ten built from exact table entries; every step is one IEEE
 * double operation, so host and device agree bit for bit.  Decimal strings within a few ulp
 * of a threshold are outside the defined behaviour (Perl's own atof is not correctly rounded
 * there either). */
#ifndef PG_PERLNUM_H
#define PG_PERLNUM_H

#ifdef __CUDACC__
#define PG_HD __host__ __device__
#else
#define PG_HD
#endif

PG_HD static inline int pg_pn_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
PG_HD static inline int pg_pn_lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

PG_HD static inline double pg_pn_mul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
PG_HD static inline double pg_pn_div(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}

PG_HD static inline double pg_pn_pow10(int e)      /* e >= 0; exact up to 10^22 */
{
    const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                          1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    double r = 1.0;
    while (e > 22) {
        r = pg_pn_mul(r, 1e22);
        e -= 22;
        if (r > 1e300) return r * 1e300;           /* overflow to inf */
    }
    return pg_pn_mul(r, t[e]);
}

PG_HD static inline double pg_perl_number(const char *s, int n)
{
    int p = 0;
    while (p < n && pg_pn_space((unsigned char)s[p])) p++;
    int neg = 0;
    if (p < n && (s[p] == '-' || s[p] == '+')) { neg = s[p] == '-'; p++; }
    /* inf / infinity / nan */
    if (n - p >= 3) {
        const int a = pg_pn_lower((unsigned char)s[p]), b = pg_pn_lower((unsigned char)s[p + 1]), c = pg_pn_lower((unsigned char)s[p + 2]);
        if (a == 'i' && b == 'n' && c == 'f') {
            const double inf = 1e300 * 1e300;
            return neg ? -inf : inf;
        }
        if (a == 'n' && b == 'a' && c == 'n') {
            const double inf = 1e300 * 1e300;
            return inf - inf;
        }
    }
    unsigned long long m = 0;
    int nd = 0, dexp = 0, any = 0, seen_point = 0;
    for (; p < n; p++) {
        const int c = (unsigned char)s[p];
        if (c >= '0' && c <= '9') {
            any = 1;
            if (nd < 19) {
                m = m * 10ULL + (unsigned long long)(c - '0');
                if (m) nd++;
                if (seen_point) dexp--;
            } else if (!seen_point) {
                dexp++;                             /* dropped integer digit keeps its weight */
            }
        } else if (c == '.' && !seen_point) {
            seen_point = 1;
        } else {
            break;
        }
    }
    if (!any) return 0.0;
    if (p < n && (s[p] == 'e' || s[p] == 'E')) {
        int q = p + 1, eneg = 0;
        if (q < n && (s[q] == '-' || s[q] == '+')) { eneg = s[q] == '-'; q++; }
        if (q < n && s[q] >= '0' && s[q] <= '9') {
            int ev = 0;
            for (; q < n && s[q] >= '0' && s[q] <= '9'; q++)
                if (ev < 100000) ev = ev * 10 + (s[q] - '0');
            dexp += eneg ? -ev : ev;
        }
    }
    double r = (double)m;
    if (m != 0) {
        if (dexp > 0) {
            r = dexp > 400 ? r * 1e300 * 1e300 : pg_pn_mul(r, pg_pn_pow10(dexp));
        } else if (dexp < 0) {
            int d = -dexp;
            if (d > 800) d = 800;
            while (d > 300) { r = pg_pn_div(r, 1e300); d -= 300; }   /* keep the divisor finite */
            r = pg_pn_div(r, pg_pn_pow10(d));
        }
    }
    return neg ? -r : r;
}

#endif
