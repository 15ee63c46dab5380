/* oracle/cpu/curve_tmpl.h -- short-Weierstrass (a = 0) Jacobian arithmetic and a
 * plain Pippenger MSM, instantiated twice (G1 over Fq, G2 over Fq2).
 * Define before including:  CT_F (field element type), CT_(name) (field op prefix
 * macro), CT_P(name) (point symbol prefix), CT_BYTES (affine coordinate bytes).
 * Formulas: EFD dbl-2009-l, add-2007-bl, madd-2007-bl (public). */

typedef struct { CT_F x, y; int inf; } CT_P(aff);
typedef struct { CT_F x, y, z; } CT_P(jac);   /* z == 0 <=> infinity */

static inline void CT_P(jac_set_inf)(CT_P(jac) *r) { CT_(one)(&r->x); CT_(one)(&r->y); CT_(zero)(&r->z); }
static inline int CT_P(jac_is_inf)(const CT_P(jac) *p) { return CT_(is_zero)(&p->z); }

static inline void CT_P(jac_dbl)(CT_P(jac) *r, const CT_P(jac) *p) {
    if (CT_P(jac_is_inf)(p)) { *r = *p; return; }
    CT_F A, B, C, D, E, Fv, t, X3, Y3, Z3;
    CT_(sqr)(&A, &p->x);
    CT_(sqr)(&B, &p->y);
    CT_(sqr)(&C, &B);
    CT_(add)(&t, &p->x, &B);
    CT_(sqr)(&t, &t);
    CT_(sub)(&t, &t, &A);
    CT_(sub)(&t, &t, &C);
    CT_(dbl)(&D, &t);
    CT_(dbl)(&E, &A);
    CT_(add)(&E, &E, &A);
    CT_(sqr)(&Fv, &E);
    CT_(dbl)(&t, &D);
    CT_(sub)(&X3, &Fv, &t);
    CT_(sub)(&t, &D, &X3);
    CT_(mul)(&Y3, &E, &t);
    CT_(dbl)(&t, &C); CT_(dbl)(&t, &t); CT_(dbl)(&t, &t);
    CT_(sub)(&Y3, &Y3, &t);
    CT_(mul)(&Z3, &p->y, &p->z);
    CT_(dbl)(&Z3, &Z3);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static inline void CT_P(jac_madd)(CT_P(jac) *r, const CT_P(jac) *p, const CT_P(aff) *q) {
    if (q->inf) { *r = *p; return; }
    if (CT_P(jac_is_inf)(p)) { r->x = q->x; r->y = q->y; CT_(one)(&r->z); return; }
    CT_F Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
    CT_(sqr)(&Z1Z1, &p->z);
    CT_(mul)(&U2, &q->x, &Z1Z1);
    CT_(mul)(&S2, &q->y, &p->z);
    CT_(mul)(&S2, &S2, &Z1Z1);
    if (CT_(eq)(&U2, &p->x)) {
        if (CT_(eq)(&S2, &p->y)) { CT_P(jac_dbl)(r, p); return; }
        CT_P(jac_set_inf)(r); return;
    }
    CT_(sub)(&H, &U2, &p->x);
    CT_(sqr)(&HH, &H);
    CT_(dbl)(&I, &HH); CT_(dbl)(&I, &I);
    CT_(mul)(&J, &H, &I);
    CT_(sub)(&rr, &S2, &p->y);
    CT_(dbl)(&rr, &rr);
    CT_(mul)(&V, &p->x, &I);
    CT_(sqr)(&X3, &rr);
    CT_(sub)(&X3, &X3, &J);
    CT_(dbl)(&t, &V);
    CT_(sub)(&X3, &X3, &t);
    CT_(sub)(&t, &V, &X3);
    CT_(mul)(&Y3, &rr, &t);
    CT_(mul)(&t, &p->y, &J);
    CT_(dbl)(&t, &t);
    CT_(sub)(&Y3, &Y3, &t);
    CT_(add)(&Z3, &p->z, &H);
    CT_(sqr)(&Z3, &Z3);
    CT_(sub)(&Z3, &Z3, &Z1Z1);
    CT_(sub)(&Z3, &Z3, &HH);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static inline void CT_P(jac_add)(CT_P(jac) *r, const CT_P(jac) *p, const CT_P(jac) *q) {
    if (CT_P(jac_is_inf)(p)) { *r = *q; return; }
    if (CT_P(jac_is_inf)(q)) { *r = *p; return; }
    CT_F Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
    CT_(sqr)(&Z1Z1, &p->z);
    CT_(sqr)(&Z2Z2, &q->z);
    CT_(mul)(&U1, &p->x, &Z2Z2);
    CT_(mul)(&U2, &q->x, &Z1Z1);
    CT_(mul)(&S1, &p->y, &q->z); CT_(mul)(&S1, &S1, &Z2Z2);
    CT_(mul)(&S2, &q->y, &p->z); CT_(mul)(&S2, &S2, &Z1Z1);
    if (CT_(eq)(&U1, &U2)) {
        if (CT_(eq)(&S1, &S2)) { CT_P(jac_dbl)(r, p); return; }
        CT_P(jac_set_inf)(r); return;
    }
    CT_(sub)(&H, &U2, &U1);
    CT_(dbl)(&I, &H); CT_(sqr)(&I, &I);
    CT_(mul)(&J, &H, &I);
    CT_(sub)(&rr, &S2, &S1); CT_(dbl)(&rr, &rr);
    CT_(mul)(&V, &U1, &I);
    CT_(sqr)(&X3, &rr);
    CT_(sub)(&X3, &X3, &J);
    CT_(dbl)(&t, &V);
    CT_(sub)(&X3, &X3, &t);
    CT_(sub)(&t, &V, &X3);
    CT_(mul)(&Y3, &rr, &t);
    CT_(mul)(&t, &S1, &J); CT_(dbl)(&t, &t);
    CT_(sub)(&Y3, &Y3, &t);
    CT_(add)(&Z3, &p->z, &q->z);
    CT_(sqr)(&Z3, &Z3);
    CT_(sub)(&Z3, &Z3, &Z1Z1);
    CT_(sub)(&Z3, &Z3, &Z2Z2);
    CT_(mul)(&Z3, &Z3, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static inline void CT_P(jac_to_aff)(CT_P(aff) *r, const CT_P(jac) *p) {
    if (CT_P(jac_is_inf)(p)) { CT_(zero)(&r->x); CT_(zero)(&r->y); r->inf = 1; return; }
    CT_F zi, zi2;
    CT_(inv)(&zi, &p->z);
    CT_(sqr)(&zi2, &zi);
    CT_(mul)(&r->x, &p->x, &zi2);
    CT_(mul)(&zi2, &zi2, &zi);
    CT_(mul)(&r->y, &p->y, &zi2);
    r->inf = 0;
}

/* boundary format: x || y canonical little-endian; all-zero = infinity */
static inline int CT_P(aff_from_bytes)(CT_P(aff) *r, const uint8_t *b) {
    int allz = 1;
    for (int i = 0; i < 2 * CT_BYTES; i++) if (b[i]) { allz = 0; break; }
    if (allz) { CT_(zero)(&r->x); CT_(zero)(&r->y); r->inf = 1; return 0; }
    r->inf = 0;
    return CT_(from_bytes)(&r->x, b) | CT_(from_bytes)(&r->y, b + CT_BYTES);
}
static inline void CT_P(aff_to_bytes)(uint8_t *b, const CT_P(aff) *p) {
    if (p->inf) { memset(b, 0, 2 * CT_BYTES); return; }
    CT_(to_bytes)(b, &p->x);
    CT_(to_bytes)(b + CT_BYTES, &p->y);
}

/* scalars are canonical 4x64 integers (NOT Montgomery) */
static inline uint32_t CT_P(window)(const uint64_t s[4], int bit, int c) {
    int w = bit >> 6, o = bit & 63;
    uint64_t v = s[w] >> o;
    if (o + c > 64 && w < 3) v |= s[w + 1] << (64 - o);
    return (uint32_t)(v & ((1ULL << c) - 1));
}

static void CT_P(scalar_mul)(CT_P(jac) *r, const CT_P(aff) *p, const uint64_t s[4]) {
    CT_P(jac) acc; CT_P(jac_set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        CT_P(jac_dbl)(&acc, &acc);
        if ((s[i >> 6] >> (i & 63)) & 1) CT_P(jac_madd)(&acc, &acc, p);
    }
    *r = acc;
}

static int CT_P(msm_window_bits)(size_t n) {
    int c = 1; while ((1ULL << (c + 1)) <= n) c++;     /* floor(log2 n) */
    c = c > 4 ? c - 2 : 2;
    return c > 16 ? 16 : c;
}

/* Pippenger: per-window bucket accumulation, running-sum reduction, Horner combine.
 * Windows run in parallel under OpenMP when called from a serial region. */
static void CT_P(msm)(CT_P(jac) *out, const CT_P(aff) *pts, const uint64_t (*sc)[4], size_t n) {
    CT_P(jac) res; CT_P(jac_set_inf)(&res);
    if (n == 0) { *out = res; return; }
    int c = CT_P(msm_window_bits)(n);
    int nw = (254 + c - 1) / c;
    size_t nb = ((size_t)1 << c) - 1;
    CT_P(jac) *wsum = (CT_P(jac) *)malloc(sizeof(CT_P(jac)) * nw);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nw; w++) {
        CT_P(jac) *bk = (CT_P(jac) *)malloc(sizeof(CT_P(jac)) * nb);
        for (size_t b = 0; b < nb; b++) CT_P(jac_set_inf)(&bk[b]);
        for (size_t i = 0; i < n; i++) {
            uint32_t d = CT_P(window)(sc[i], w * c, c);
            if (d) CT_P(jac_madd)(&bk[d - 1], &bk[d - 1], &pts[i]);
        }
        CT_P(jac) run, sum; CT_P(jac_set_inf)(&run); CT_P(jac_set_inf)(&sum);
        for (size_t b = nb; b-- > 0;) {
            CT_P(jac_add)(&run, &run, &bk[b]);
            CT_P(jac_add)(&sum, &sum, &run);
        }
        wsum[w] = sum;
        free(bk);
    }
    for (int w = nw - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) CT_P(jac_dbl)(&res, &res);
        CT_P(jac_add)(&res, &res, &wsum[w]);
    }
    free(wsum);
    *out = res;
}
