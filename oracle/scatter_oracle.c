/* TEST INFRASTRUCTURE -- CPU restatement of the Monte Carlo scatter transport (kernel 3 of BASELINE.json's north_star).
 *
 * The reference has no transport kernel any more (deepdrr/projector/projector.py:530-531 raises DeprecationError); what it ships
 * are MC-GPU's material tables and two host helpers:
 *   - mean free paths / Rayleigh cumulative maxima   deepdrr/projector/mcgpu_mfp_data.py:39-45
 *   - RITA sampling of the squared form factor       deepdrr/projector/rita.py:129-183 (sample_rita: binary search + rational inverse)
 *   - Compton shell data FCO, UICO, FJ0              deepdrr/projector/mcgpu_compton_data.py:122-166
 *   - detector-plane intersection and bounds test    deepdrr/projector/plane_surface.py:44-110
 * SURVEY.md App. C (i) asks for "a CPU restatement with the same tables and a counter-based RNG": this file restates the
 * published MC-GPU scheme (Badal & Badano, Med. Phys. 36, 2009; PENELOPE-2006 GRAa / GCOa for the two scattering processes) in
 * plain C with the SAME Philox4x32-10 stream per photon id as csrc/drr_scatter.cu (cuRAND's Philox: key = seed, counter high
 * words = photon id, 4 outputs per counter, uniform = x * 2^-32 + 2^-33), so that a photon id names the same history on both
 * sides up to the last-ulp differences of libm against the CUDA math library.  Parity with the reference is UNPINNED (there is
 * nothing to pin to); this restatement pins the CUDA kernel to an independent implementation of the same published algorithm.
 * Only tests/ may load it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_mat, n_e;
    float e0, de;
    const float* mfp;          /* [n_mat][n_e][5]: Rayleigh, Compton, photoelectric, total mean free path (mm), Rayleigh max cumul. prob */
    const float* rita;         /* [n_mat][128][4]: x^2, P, A, B */
    const float* compton;      /* [n_mat][30][3]: electrons, ionisation energy (eV), J0 m_e c */
    const int* nshell;         /* [n_mat] */
    const float* inv_rho_nom;  /* [n_mat] */
    const float* majorant;     /* [n_e] */
    const int* mat_of_label;   /* [M] */
    const float* s0;           /* [n_mat][n_e]: S(E, theta = pi) x 1.001, the normalisation of the Compton rejection */
    int V;                     /* volumes; a point inside several belongs to the one with the smallest priority value */
    int priority[8], enabled[8];
    const float* dens[8];      /* [ni][nj][nk] g/cm^3 (NumPy order of Volume.data) */
    const uint8_t* lab[8];     /* [ni][nj][nk] global material index */
    int shape[8][3];
    float ijk[8][12];          /* ijk_from_world per volume */
    float p_idx[12], w2i[9], src[3];
    int W, H, n_bins;
    const float* spec_e_keV;
    const float* spec_cdf;
} sc_scene;

/* ---- Philox4x32-10 as cuRAND drives it (curand_init(seed, subsequence, 0); curand_uniform) ---------------------------- */
typedef struct { uint32_t ctr[4], key[2], out[4]; int pos; } philox;

static void philox_block(const uint32_t c[4], const uint32_t k[2], uint32_t o[4]) {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
static void philox_init(philox* s, uint64_t seed, uint64_t subsequence) {
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = s->ctr[1] = 0; s->ctr[2] = (uint32_t)subsequence; s->ctr[3] = (uint32_t)(subsequence >> 32);
    s->pos = 0;
    philox_block(s->ctr, s->key, s->out);
}
static float philox_uniform(philox* s) {
    uint32_t x = s->out[s->pos++];
    if (s->pos == 4) {
        if (++s->ctr[0] == 0 && ++s->ctr[1] == 0 && ++s->ctr[2] == 0) ++s->ctr[3];
        philox_block(s->ctr, s->key, s->out);
        s->pos = 0;
    }
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);  /* (0, 1] */
}

/* ---- interaction tables ------------------------------------------------------------------------------------------------ */
static void mfp_lookup(const sc_scene* S, int mat, float E, float* iray, float* ico, float* itot, float* pmax) {
    float f = (E - S->e0) / S->de;
    int i = (int)f;
    if (i < 0) i = 0;
    if (i > S->n_e - 2) i = S->n_e - 2;
    float w = fminf(fmaxf(f - (float)i, 0.0f), 1.0f);
    const float* a = S->mfp + ((size_t)mat * S->n_e + i) * 5;
    const float* b = a + 5;
    *iray = 1.0f / (a[0] + w * (b[0] - a[0]));
    *ico = 1.0f / (a[1] + w * (b[1] - a[1]));
    *itot = 1.0f / (a[3] + w * (b[3] - a[3]));
    *pmax = a[4] + w * (b[4] - a[4]);
}

/* new direction after a polar deflection cos(theta) = cost and azimuth phi (PENELOPE's DIRECT) */
static void rotate_dir(float* dx, float* dy, float* dz, float cost, float phi) {
    float sint = sqrtf(fmaxf(0.0f, 1.0f - cost * cost));
    float sp = sinf(phi), cp = cosf(phi);
    float x = *dx, y = *dy, z = *dz;
    float dxy = x * x + y * y;
    if (dxy > 1e-10f) {
        float s = sqrtf(dxy);
        float nx = x * cost + sint * (x * z * cp - y * sp) / s;
        float ny = y * cost + sint * (y * z * cp + x * sp) / s;
        float nz = z * cost - s * sint * cp;
        x = nx; y = ny; z = nz;
    } else {
        float sgn = z > 0 ? 1.0f : -1.0f;
        x = sint * cp; y = sint * sp; z = sgn * cost;
    }
    float n = 1.0f / sqrtf(x * x + y * y + z * z);
    *dx = x * n; *dy = y * n; *dz = z * n;
}

/* Rayleigh: x^2 from the RITA table of the squared form factor (rita.py:129-183), (1 + cos^2) / 2 rejection (GRAa) */
static float sample_rayleigh(const sc_scene* S, int mat, float E, float pmax, philox* st) {
    const float* R = S->rita + (size_t)mat * 128 * 4;
    float xmax = E * 8.065535669099010e-5f;
    float x2max = fminf(xmax * xmax, R[127 * 4]);
    float cost = 1.0f;
    if (xmax < 1e-4f) {
        do { cost = 1.0f - 2.0f * philox_uniform(st); } while (philox_uniform(st) > 0.5f * (1.0f + cost * cost));
        return cost;
    }
    for (int tries = 0; tries < 64; tries++) {
        float ru = philox_uniform(st) * pmax;
        int lo = 0, hi = 127;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ru > R[mid * 4 + 1]) lo = mid; else hi = mid; }
        float rr = ru - R[lo * 4 + 1], x2;
        if (rr > 1e-16f) {
            float d = R[hi * 4 + 1] - R[lo * 4 + 1], a = R[lo * 4 + 2], b = R[lo * 4 + 3];
            x2 = R[lo * 4] + ((1.0f + a + b) * d * rr / (d * d + (a * d + b * rr) * rr)) * (R[hi * 4] - R[lo * 4]);
        } else x2 = R[lo * 4];
        cost = 1.0f - 2.0f * x2 / x2max;
        cost = fmaxf(-1.0f, fminf(1.0f, cost));
        if (philox_uniform(st) <= 0.5f * (1.0f + cost * cost)) break;
    }
    return cost;
}

/* Compton: impulse approximation with analytical one-electron profiles (PENELOPE-2006 GCOa; shell data of
 * mcgpu_compton_data.py:122-166).  n_i(p) = 1/2 exp(1/2 - (d1 - d2 J p)^2) for p < 0, 1 - 1/2 exp(1/2 - (d1 + d2 J p)^2) else. */
static float profile_cdf(const float* C, int i, float E, float cdt1) {
    const float REV = 510998.918f, D2 = 1.4142135623731f, D1 = 0.70710678118655f;
    float U = C[3 * i + 1];
    float aux = E * (E - U) * cdt1;
    float pz = C[3 * i + 2] * (aux - REV * U) / (REV * sqrtf(aux + aux + U * U));
    float q = pz > 0.0f ? D1 + D2 * pz : D1 - D2 * pz;
    float h = 0.5f * expf(0.5f - q * q);
    return pz > 0.0f ? 1.0f - h : h;
}
static float sample_compton(const sc_scene* S, int mat, float* Eio, philox* st) {
    const float REV = 510998.918f, D2 = 1.4142135623731f, D1 = 0.70710678118655f, D12 = 0.5f;
    float E = *Eio;
    float ek = E / REV, ek2 = ek + ek + 1.0f, eks = ek * ek, ek1 = eks - ek2 - 1.0f;
    float taumin = 1.0f / ek2, taum2 = taumin * taumin;
    float a1 = logf(ek2), a2 = a1 + 2.0f * ek * (1.0f + ek) * taum2;
    const float* C = S->compton + (size_t)mat * 30 * 3;
    int ns = S->nshell[mat];
    float s0;  /* S(E, pi) from the table (compton_samples: computed here) */
    if (S->s0) {
        float f = (E - S->e0) / S->de;
        int i = (int)f;
        if (i < 0) i = 0;
        if (i > S->n_e - 2) i = S->n_e - 2;
        float w = fminf(fmaxf(f - (float)i, 0.0f), 1.0f);
        const float* a = S->s0 + (size_t)mat * S->n_e + i;
        s0 = a[0] + w * (a[1] - a[0]);
    } else {
        s0 = 0.0f;
        for (int i = 0; i < ns; i++)
            if (C[3 * i + 1] < E) s0 += C[3 * i] * profile_cdf(C, i, E, 2.0f);
    }
    float rn[30], pac[30], tau = 1.0f, cdt1 = 0.0f, sfun = 0.0f;
    for (int tries = 0; tries < 200; tries++) {
        if (philox_uniform(st) * a2 < a1) tau = powf(taumin, philox_uniform(st));
        else tau = sqrtf(1.0f + philox_uniform(st) * (taum2 - 1.0f));
        cdt1 = (1.0f - tau) / (ek * tau);
        sfun = 0.0f;
        for (int i = 0; i < ns; i++) {
            if (C[3 * i + 1] < E) { rn[i] = profile_cdf(C, i, E, cdt1); sfun += C[3 * i] * rn[i]; pac[i] = sfun; }
            else { rn[i] = 0.0f; pac[i] = sfun - 1.0e-6f; }
        }
        float tst = sfun * (1.0f + tau * (ek1 + tau * (ek2 + tau * eks))) / (eks * tau * (1.0f + tau * tau));
        if (!(philox_uniform(st) * s0 > tst)) break;
    }
    float cdt = 1.0f - cdt1;
    if (!(sfun > 0.0f)) { *Eio = E * tau; return fmaxf(-1.0f, fminf(1.0f, cdt)); }
    float pz = 0.0f;
    for (int tries = 0; tries < 200; tries++) {
        float tst = sfun * philox_uniform(st);
        int ish = ns - 1;
        for (int i = 0; i < ns; i++) if (pac[i] > tst) { ish = i; break; }
        float a = philox_uniform(st) * rn[ish];
        if (a < 0.5f) pz = (D1 - sqrtf(D12 - logf(a + a))) / (D2 * C[3 * ish + 2]);
        else pz = (sqrtf(D12 - logf(2.0f - a - a)) - D1) / (D2 * C[3 * ish + 2]);
        if (pz < -1.0f) continue;
        float xqc = 1.0f + tau * (tau - 2.0f * cdt);
        float af = sqrtf(xqc) * (1.0f + tau * (tau - cdt) / xqc);
        float fmax_ = af > 0.0f ? 1.0f + af * 0.2f : 1.0f - af * 0.2f;
        float fpz = 1.0f + af * fmaxf(fminf(pz, 0.2f), -0.2f);
        if (!(philox_uniform(st) * fmax_ > fpz)) break;
    }
    float t = pz * pz, b1 = 1.0f - t * tau * tau, b2 = 1.0f - t * tau * cdt;
    float root = sqrtf(fabsf(b2 * b2 - b1 * (1.0f - t)));
    *Eio = E * (tau / b1) * (pz > 0.0f ? b2 + root : b2 - root);
    return fmaxf(-1.0f, fminf(1.0f, cdt));
}

/* [t0, t1] along (x, d) in which the photon can be inside some volume: union of the slab intervals of the volumes it hits */
static void span(const sc_scene* S, float x, float y, float z, float dx, float dy, float dz, float* t0, float* t1) {
    *t0 = INFINITY; *t1 = -INFINITY;
    for (int v = 0; v < S->V; v++) {
        if (!S->enabled[v]) continue;
        const float* A = S->ijk[v];
        const float d[3] = {A[0] * dx + A[1] * dy + A[2] * dz, A[4] * dx + A[5] * dy + A[6] * dz, A[8] * dx + A[9] * dy + A[10] * dz};
        const float p[3] = {A[0] * x + A[1] * y + A[2] * z + A[3], A[4] * x + A[5] * y + A[6] * z + A[7], A[8] * x + A[9] * y + A[10] * z + A[11]};
        const float mx[3] = {(float)S->shape[v][0] - 0.5f, (float)S->shape[v][1] - 0.5f, (float)S->shape[v][2] - 0.5f};
        float a0 = 0.0f, a1 = INFINITY;
        int miss = 0;
        for (int a = 0; a < 3; a++) {
            if (d[a] != 0.0f) {
                float ta = (-0.5f - p[a]) / d[a], tb = (mx[a] - p[a]) / d[a];
                a0 = fmaxf(a0, fminf(ta, tb)); a1 = fminf(a1, fmaxf(ta, tb));
            } else if (p[a] < -0.5f || p[a] > mx[a]) miss = 1;
        }
        if (!miss && a0 < a1) { *t0 = fminf(*t0, a0); *t1 = fmaxf(*t1, a1); }
    }
}

/* Photon ids [offset, offset + n): tally[H*W] in 2^-16 eV fixed point, counters[8] as in csrc/drr_scatter.cu. */
int drr_scatter_oracle(const sc_scene* S, uint64_t n_photons, uint64_t offset, uint64_t seed, uint64_t* tally, double* counters) {
    for (uint64_t id = 0; id < n_photons; id++) {
        philox st;
        philox_init(&st, seed, offset + id);
        /* source: energy from the spectrum CDF, direction uniform over the detector area, weight = cos^3 ~ |r|^-3 */
        float xi = philox_uniform(&st);
        int lo = 0, hi = S->n_bins - 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (S->spec_cdf[mid] < xi) lo = mid + 1; else hi = mid; }
        float E = S->spec_e_keV[lo] * 1000.0f;
        float u = philox_uniform(&st) * S->W, v = philox_uniform(&st) * S->H;
        const float* w = S->w2i;
        float dx = u * w[0] + v * w[1] + w[2], dy = u * w[3] + v * w[4] + w[5], dz = u * w[6] + v * w[7] + w[8];
        float rl = sqrtf(dx * dx + dy * dy + dz * dz);
        dx /= rl; dy /= rl; dz /= rl;
        float wgt = 1.0f / (rl * rl * rl);
        float x = S->src[0], y = S->src[1], z = S->src[2];
        counters[0] += (double)E * wgt;
        float t0, t1;
        span(S, x, y, z, dx, dy, dz, &t0, &t1);
        if (!(t0 < t1)) { counters[1] += (double)E * wgt; continue; }
        float t = t0 + 1e-4f;
        int n_scat = 0, alive = 1;
        /* Woodcock tracking with the per-energy majorant */
        for (int guard = 0; guard < 100000 && alive; guard++) {
            float f = (E - S->e0) / S->de;
            int ie = (int)f;
            if (ie < 0) ie = 0;
            if (ie > S->n_e - 2) ie = S->n_e - 2;
            float wq = fminf(fmaxf(f - (float)ie, 0.0f), 1.0f);
            float smax = S->majorant[ie] + wq * (S->majorant[ie + 1] - S->majorant[ie]);
            smax *= 1.0001f;
            t += -logf(philox_uniform(&st)) / smax;
            if (t > t1) break;
            const float X = x + t * dx, Y = y + t * dy, Z = z + t * dz;
            int best = -1, best_pr = 0x7fffffff;
            size_t o = 0;
            for (int vv = 0; vv < S->V; vv++) {
                if (!S->enabled[vv] || S->priority[vv] >= best_pr) continue;
                const float* A = S->ijk[vv];
                const float qi = A[0] * X + A[1] * Y + A[2] * Z + A[3], qj = A[4] * X + A[5] * Y + A[6] * Z + A[7], qk = A[8] * X + A[9] * Y + A[10] * Z + A[11];
                const int ni = S->shape[vv][0], nj = S->shape[vv][1], nk = S->shape[vv][2];
                if (qi < -0.5f || qi > (float)ni - 0.5f || qj < -0.5f || qj > (float)nj - 0.5f || qk < -0.5f || qk > (float)nk - 0.5f) continue;
                int vi = (int)floorf(qi + 0.5f), vj = (int)floorf(qj + 0.5f), vk = (int)floorf(qk + 0.5f);
                vi = vi < 0 ? 0 : (vi > ni - 1 ? ni - 1 : vi);
                vj = vj < 0 ? 0 : (vj > nj - 1 ? nj - 1 : vj);
                vk = vk < 0 ? 0 : (vk > nk - 1 ? nk - 1 : vk);
                best = vv; best_pr = S->priority[vv];
                o = ((size_t)vi * nj + vj) * nk + vk;
            }
            if (best < 0) continue;  /* vacuum between the volumes */
            int mat = S->mat_of_label[S->lab[best][o]];
            float rho = S->dens[best][o];
            float iray, ico, itot, pmax;
            mfp_lookup(S, mat, E, &iray, &ico, &itot, &pmax);
            float scale = rho * S->inv_rho_nom[mat];
            if (philox_uniform(&st) * smax >= itot * scale) continue;  /* virtual interaction */
            float r = philox_uniform(&st) * itot;
            x = X; y = Y; z = Z;
            float cost;
            if (r < iray) { cost = sample_rayleigh(S, mat, E, pmax, &st); counters[6] += 1.0; }
            else if (r < iray + ico) { float E0 = E; cost = sample_compton(S, mat, &E, &st); counters[2] += (double)(E0 - E) * wgt; counters[7] += 1.0; }
            else { counters[2] += (double)E * wgt; alive = 0; break; }
            if (E < S->e0) { counters[2] += (double)E * wgt; alive = 0; break; }
            rotate_dir(&dx, &dy, &dz, cost, 6.283185307f * philox_uniform(&st));
            n_scat++;
            span(S, x, y, z, dx, dy, dz, &t0, &t1);
            t = 0.0f;
        }
        if (!alive) continue;
        if (n_scat == 0) { counters[3] += (double)E * wgt; continue; }
        /* detector plane (plane_surface.py:44-110): the plane w == 1 of the scaled projection matrix, bounds test in pixels */
        const float* P = S->p_idx;
        float w0 = P[8] * x + P[9] * y + P[10] * z + P[11];
        float wd = P[8] * dx + P[9] * dy + P[10] * dz;
        int hit = 0;
        if (wd > 1e-9f) {
            float s = (1.0f - w0) / wd;
            if (s > 0.0f) {
                float Xd = x + s * dx, Yd = y + s * dy, Zd = z + s * dz;
                float uu = P[0] * Xd + P[1] * Yd + P[2] * Zd + P[3];
                float vv = P[4] * Xd + P[5] * Yd + P[6] * Zd + P[7];
                int iu = (int)floorf(uu), iv = (int)floorf(vv);
                if (iu >= 0 && iu < S->W && iv >= 0 && iv < S->H) {
                    hit = 1;
                    tally[(size_t)iv * S->W + iu] += (uint64_t)((double)E * (double)wgt * 65536.0 + 0.5);
                }
            }
        }
        counters[hit ? 4 : 5] += (double)E * wgt;
    }
    return 0;
}

/* n Compton events of a photon of energy E in table material `mat` (unit test of the sampler): cos(theta) and E' per event. */
int drr_compton_oracle(const sc_scene* S, int mat, float E, uint64_t seed, int n, float* cost, float* e_out) {
    for (int i = 0; i < n; i++) {
        philox st;
        philox_init(&st, seed, (uint64_t)i);
        float e = E;
        cost[i] = sample_compton(S, mat, &e, &st);
        e_out[i] = e;
    }
    return 0;
}

size_t drr_scatter_oracle_scene_size(void) { return sizeof(sc_scene); }
