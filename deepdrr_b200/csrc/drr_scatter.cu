// Monte Carlo scatter photon transport (libdrr_b200, sm_100a): kernel (3) of BASELINE.json's north_star.
//
// The reference removed its scatter kernel: `Projector(scatter_num > 0)` raises DeprecationError
// (deepdrr/projector/projector.py:530-531) and only the MC-GPU data tables + a Python RITA sampler remain
// (mcgpu_mfp_data.py, mcgpu_rita_samplers.py, mcgpu_compton_data.py, rita.py:129-183, plane_surface.py:44-110).
// There is nothing to be bit-compatible with (parity unpinned); this kernel follows the
// published MC-GPU scheme (Badal & Badano, Med. Phys. 36, 2009) on those tables:
//   * photon energy from the spectrum CDF, direction uniform over the detector area with the solid-angle
//     weight cos^3(theta) carried as a statistical weight;
//   * Woodcock (delta) tracking through the voxel volumes (the one with the smallest priority value owns a point that lies in
//     several, as in the ray march; between volumes is vacuum) with the per-energy minimum total mean free path;
//   * interaction type by the ratio of inverse mean free paths; photoelectric absorption ends the history;
//   * Rayleigh: angle from the RITA-sampled squared form factor (x^2 tables) with (1 + cos^2)/2 rejection;
//   * Compton: relativistic impulse approximation with the analytical one-electron profiles of the shipped shell tables
//     (binding effects and Doppler broadening, PENELOPE's GCOa as in MC-GPU);
//   * only photons that scattered at least once are tallied (the primary comes from the ray march):
//     energy x weight into the pixel the photon hits (detector plane = image plane of the camera).
// Tallies are 64-bit fixed point (2^-16 eV), so the sum is exact and independent of the order in which
// threads, launches or GPUs add their photons: any split of the photon range gives bit-identical tallies.
// Organisation: a warp keeps a pool of photon records in shared memory and works on the records that wait for the same thing
// (tracking steps, a Rayleigh sample, a Compton try) side by side -- see scatter_kernel.  The random numbers are cuRAND's
// Philox4x32-10 stream per photon id, restated below.
#include <math_constants.h>

#include "drr_device.cuh"

struct ScatterTables {
    int n_mat, n_e;
    float e0, de;                // energy grid (eV)
    const float* mfp;            // [n_mat][n_e][5]: Rayleigh, Compton, photoelectric, total (mm at nominal density), Rayleigh max cumul. prob
    const float* rita;           // [n_mat][128][4]: x^2, P, A, B
    const float* compton;        // [n_mat][30][3]: electrons, ionisation energy (eV), J0
    const int* nshell;           // [n_mat]
    const float* inv_rho_nom;    // [n_mat]
    const float* majorant;       // [n_e]: max over materials of rho_max / rho_nom / mfp_total  (1/mm)
    const int* mat_of_label;     // [M] global material index -> table material
    const float* s0;             // [n_mat][n_e]: incoherent scattering function at theta = pi (its maximum), x 1.001
};

struct ScatterParams {
    ScatterTables T;
    int V;                       // volumes of the scene; a point inside several belongs to the one with the smallest priority value,
    int priority[DRR_MAX_VOLUMES];  // as in the ray march (K.cu:458-496); outside all of them is vacuum
    int enabled[DRR_MAX_VOLUMES];
    VolDev vol[DRR_MAX_VOLUMES];
    float ijk[DRR_MAX_VOLUMES][12];  // ijk_from_world per volume
    float p_idx[12];             // index_from_world (3x4): (u*w, v*w, w) = P (x, 1)
    float w2i[9];
    float src[3];
    int W, H;
    int n_bins;
    const float* spec_e_keV;
    const float* spec_cdf;       // [n_bins] cumulative probability of max(pdf, 0)
    unsigned long long n_photons, photon_offset, seed;
    unsigned long long* tally;   // [H*W] fixed point, 2^-16 eV
    double* counters;            // [8] energy bookkeeping (eV * weight): 0 emitted, 1 missed volume, 2 absorbed, 3 exit primary,
                                 //     4 exit scattered & detected, 5 exit scattered & not detected, 6 #rayleigh, 7 #compton
};

// Philox4x32-10 exactly as cuRAND drives it for curand_init(seed, subsequence, 0) + curand_uniform: key = seed, counter =
// (block, 0, subsequence lo, subsequence hi), four outputs per block handed out in order, uniform = x * 2^-32 + 2^-33 in (0, 1].  Restated here: 8 words of state per photon instead of cuRAND's
// struct, and the ten rounds out of line (they were inlined at a dozen call sites, 46 instructions each, in a kernel whose code did
// not fit the instruction cache).  The block counter is 32 bits: a history would need 2^34 draws to wrap it.
// (Tried: carrying the next block as well and computing it only where the warp is convergent, so that the rounds run with half
// the lanes active instead of four.  5 % slower: the kernel waits for memory, not for issue slots, and the records grow.)
struct Philox {
    unsigned k0, k1;          // key = seed (the same for every photon)
    unsigned c0, c2, c3;      // block index; subsequence = photon id
    unsigned o0, o1, o2, o3;  // current block
    int pos;
};
__device__ __noinline__ uint4 philox_block(unsigned c0, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
    unsigned c1 = 0;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const unsigned n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ void philox_start(Philox& s, unsigned long long seed, unsigned long long subsequence) {
    s.k0 = (unsigned)seed; s.k1 = (unsigned)(seed >> 32);
    s.c0 = 0; s.c2 = (unsigned)subsequence; s.c3 = (unsigned)(subsequence >> 32);
    const uint4 o = philox_block(s.c0, s.c2, s.c3, s.k0, s.k1);
    s.o0 = o.x; s.o1 = o.y; s.o2 = o.z; s.o3 = o.w;
    s.pos = 0;
}
__device__ __forceinline__ float philox_uniform(Philox& s) {
    const unsigned x = s.pos == 0 ? s.o0 : (s.pos == 1 ? s.o1 : (s.pos == 2 ? s.o2 : s.o3));
    if (++s.pos == 4) {  // as cuRAND: the next block is computed as soon as the fourth output is taken
        s.c0 += 1;
        const uint4 o = philox_block(s.c0, s.c2, s.c3, s.k0, s.k1);
        s.o0 = o.x; s.o1 = o.y; s.o2 = o.z; s.o3 = o.w;
        s.pos = 0;
    }
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
// (Tried: computing the next block only when the fifth output is asked for, so that a draw never ends in a call while the voxel
// reads of a tracking step are in flight -- the call waits for them.  5.55e8 -> 5.1e8 photons/s: slower.)

__device__ __forceinline__ void rotate_dir(float& dx, float& dy, float& dz, float cost, float phi) {
    float sint = sqrtf(fmaxf(0.0f, 1.0f - cost * cost));
    float sp, cp;
    __sincosf(phi, &sp, &cp);
    float dxy = dx * dx + dy * dy;
    if (dxy > 1e-10f) {
        float s = sqrtf(dxy);
        float nx = dx * cost + sint * (dx * dz * cp - dy * sp) / s;
        float ny = dy * cost + sint * (dy * dz * cp + dx * sp) / s;
        float nz = dz * cost - s * sint * cp;
        dx = nx; dy = ny; dz = nz;
    } else {
        float sgn = dz > 0 ? 1.0f : -1.0f;
        dx = sint * cp; dy = sint * sp; dz = sgn * cost;
    }
    float n = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= n; dy *= n; dz *= n;
}

__device__ float sample_rayleigh(const ScatterTables& T, int mat, float E, float pmax, Philox& st) {
    const float* R = T.rita + (size_t)mat * 128 * 4;
    float xmax = E * 8.065535669099010e-5f;
    float x2max = fminf(xmax * xmax, R[127 * 4]);
    float cost;
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

// Compton scattering in the relativistic impulse approximation with analytical one-electron Compton profiles -- the model of
// PENELOPE-2006's GCOa, which MC-GPU uses and whose shell data (electrons f_i, ionisation energy U_i, J_i(0) m_e c) the reference
// ships in mcgpu_compton_data.py:122-166:
//   1. tau = E'/E of a free electron at rest from the Klein-Nishina mixture, accepted with T(tau) S(E, theta) / S(E, pi), where
//      S = sum_i f_i Theta(E - U_i) n_i(p_i,max) and n_i is the cumulative analytical profile;
//   2. the active shell with probability ~ f_i n_i(p_i,max), the electron's momentum projection p_z from that shell's profile on
//      (-inf, p_i,max), accepted with F(p_z) / F_max (Doppler broadening);
//   3. E' from the Compton line shifted by p_z.
// `compton_try` is ONE pass of the rejection loop of step 1 (the expensive part: a loop over up to 30 shells); the kernel calls it
// once per round for every lane that waits for a Compton sample, so the lanes of a warp stay together however many tries each of
// them needs.  `compton_finish` is steps 2 and 3 for an accepted try: returns cos(theta) and replaces E.
struct ComptonTry {
    float tau, cdt1, sfun;
    float rn[32], pac[32];  // up to 30 shells; read four at a time
};

__device__ __forceinline__ bool compton_try(const ScatterTables& T, int mat, float E, Philox& st, ComptonTry& c) {
    const float REV = 510998.918f, D2 = 1.4142135623731f, D1 = 0.70710678118655f, D12 = 0.5f;
    const float ek = E / REV, ek2 = ek + ek + 1.0f, eks = ek * ek, ek1 = eks - ek2 - 1.0f;
    const float taumin = 1.0f / ek2, taum2 = taumin * taumin;
    const float a1 = logf(ek2), a2 = a1 + 2.0f * ek * (1.0f + ek) * taum2;
    const float* C = T.compton + (size_t)mat * 30 * 3;
    const int ns = T.nshell[mat];
    // S(E, theta = pi), the maximum of the incoherent scattering function over the angle: tabulated on the energy grid by the host
    // (it only normalises the rejection; a third of the sampler's shell-profile evaluations otherwise)
    float s0;
    {
        const float f = (E - T.e0) / T.de;
        const int i = max(0, min((int)f, T.n_e - 2));
        const float w = fminf(fmaxf(f - (float)i, 0.0f), 1.0f);
        const float* a = T.s0 + (size_t)mat * T.n_e + i;
        s0 = a[0] + w * (a[1] - a[0]);
    }
    float tau;
    if (philox_uniform(st) * a2 < a1) tau = powf(taumin, philox_uniform(st));
    else tau = sqrtf(1.0f + philox_uniform(st) * (taum2 - 1.0f));
    const float cdt1 = (1.0f - tau) / (ek * tau);
    float sfun = 0.0f;
    for (int i = 0; i < ns; i++) {
        const float U = C[3 * i + 1];
        if (U < E) {
            // n_i(p_i,max) for 1 - cos(theta) = cdt1
            const float aux = E * (E - U) * cdt1;
            const float pz = C[3 * i + 2] * (aux - REV * U) / (REV * sqrtf(aux + aux + U * U));
            const float q = pz > 0.0f ? D1 + D2 * pz : D1 - D2 * pz;
            const float h = 0.5f * expf(D12 - q * q);
            c.rn[i] = pz > 0.0f ? 1.0f - h : h;
            sfun += C[3 * i] * c.rn[i];
            c.pac[i] = sfun;
        } else { c.rn[i] = 0.0f; c.pac[i] = sfun - 1.0e-6f; }
    }
    c.tau = tau; c.cdt1 = cdt1; c.sfun = sfun;
    const float tst = sfun * (1.0f + tau * (ek1 + tau * (ek2 + tau * eks))) / (eks * tau * (1.0f + tau * tau));
    return !(philox_uniform(st) * s0 > tst);
}

__device__ __forceinline__ float compton_finish(const ScatterTables& T, int mat, float& E, Philox& st, const ComptonTry& c) {
    const float D2 = 1.4142135623731f, D1 = 0.70710678118655f, D12 = 0.5f;
    const float* C = T.compton + (size_t)mat * 30 * 3;
    const int ns = T.nshell[mat];
    const float tau = c.tau, sfun = c.sfun, cdt = 1.0f - c.cdt1;
    if (!(sfun > 0.0f)) { E *= tau; return fmaxf(-1.0f, fminf(1.0f, cdt)); }  // no shell can be ionised: free-electron kinematics
    float pzomc = 0.0f;
    for (int tries = 0; tries < 200; tries++) {
        const float tst = sfun * philox_uniform(st);
        int ish = ns - 1;
        // first shell whose cumulative probability exceeds tst; four loads of the (local-memory) table in flight at a time
        for (int i = 0; i < ns; i += 4) {
            const float p0 = c.pac[i], p1 = c.pac[i + 1], p2 = c.pac[i + 2], p3 = c.pac[i + 3];
            const int hit = p0 > tst ? 0 : ((i + 1 < ns && p1 > tst) ? 1 : ((i + 2 < ns && p2 > tst) ? 2 : ((i + 3 < ns && p3 > tst) ? 3 : 4)));
            if (hit < 4) { ish = i + hit; break; }
        }
        const float a = philox_uniform(st) * c.rn[ish];
        if (a < 0.5f) pzomc = (D1 - sqrtf(D12 - logf(a + a))) / (D2 * C[3 * ish + 2]);
        else pzomc = (sqrtf(D12 - logf(2.0f - a - a)) - D1) / (D2 * C[3 * ish + 2]);
        if (pzomc < -1.0f) continue;
        const float xqc = 1.0f + tau * (tau - 2.0f * cdt);
        const float af = sqrtf(xqc) * (1.0f + tau * (tau - cdt) / xqc);
        const float fpzmax = af > 0.0f ? 1.0f + af * 0.2f : 1.0f - af * 0.2f;
        const float fpz = 1.0f + af * fmaxf(fminf(pzomc, 0.2f), -0.2f);
        if (!(philox_uniform(st) * fpzmax > fpz)) break;
    }
    const float t = pzomc * pzomc, b1 = 1.0f - t * tau * tau, b2 = 1.0f - t * tau * cdt;
    const float root = sqrtf(fabsf(b2 * b2 - b1 * (1.0f - t)));
    E = E * (tau / b1) * (pzomc > 0.0f ? b2 + root : b2 - root);
    return fmaxf(-1.0f, fminf(1.0f, cdt));
}

// Photons are regrouped by what they have to do next.  A history is a run of cheap tracking steps (~180 instructions, a real
// interaction every sixth step) with expensive samplers in between (a Compton try loops over up to 30 shells); written as one loop
// per photon, a warp spends most of its instructions in a sampler with one or two lanes active (ncu: 3.0 of 32 lanes active per
// executed instruction, 1.3e8 photons/s).  Here a warp owns a pool of SC_SLOTS photon records in shared memory, each tagged with
// what it waits for -- TRACK, a Rayleigh sample, (another try of) a Compton sample -- and works in visits: it picks the tag with the
// most records, the lanes load up to 32 of them, do that one thing side by side (SC_STEPS tracking steps / one sampler round) and
// put the records back under their new tags; when half the pool is empty the 32 lanes start 32 new photons together.
// The draws of a photon come from its own Philox subsequence in the order of its own history, and a warp walks the same photon
// ids as in the one-loop form, so tallies are bit-identical to it (and to any split of the photon range).
#ifndef SC_STEPS
#define SC_STEPS 4
#endif
#ifndef SC_SLOTS
#define SC_SLOTS 64      // records per warp, a multiple of 32
#endif
#ifndef SC_MIN_BLOCKS
#define SC_MIN_BLOCKS 6
#endif
#define SC_WARPS 4

struct PhotonPool {
    float f[11][SC_SLOTS];      // x, y, z, dx, dy, dz, E, wgt, t, t1, pmax
    unsigned u[10][SC_SLOTS];   // Philox: c0, id lo, id hi, o0..o3; pos | n_try << 8 | mat << 16; n_step; n_scat
    unsigned char tag[SC_SLOTS];
    unsigned char list[32];
};

__global__ void __launch_bounds__(32 * SC_WARPS, SC_MIN_BLOCKS) scatter_kernel(const __grid_constant__ ScatterParams P) {
    __shared__ PhotonPool pools[SC_WARPS];
    PhotonPool& W = pools[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const ScatterTables& T = P.T;
    double c_loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // [t0, t1] along (x, d) in which the photon can be inside some volume: union of the slab intervals of the volumes it hits
    auto span = [&](float x, float y, float z, float dx, float dy, float dz, float& t0, float& t1) {
        t0 = CUDART_INF_F; t1 = -CUDART_INF_F;
        for (int v = 0; v < P.V; v++) {
            if (!P.enabled[v]) continue;
            const float* A = P.ijk[v];
            const float d[3] = {A[0] * dx + A[1] * dy + A[2] * dz, A[4] * dx + A[5] * dy + A[6] * dz, A[8] * dx + A[9] * dy + A[10] * dz};
            const float p[3] = {A[0] * x + A[1] * y + A[2] * z + A[3], A[4] * x + A[5] * y + A[6] * z + A[7], A[8] * x + A[9] * y + A[10] * z + A[11]};
            const float mx[3] = {(float)P.vol[v].ni - 0.5f, (float)P.vol[v].nj - 0.5f, (float)P.vol[v].nk - 0.5f};
            float a0 = 0.0f, a1 = CUDART_INF_F;
            bool miss = false;
            for (int a = 0; a < 3; a++) {
                if (d[a] != 0.0f) {
                    float ta = (-0.5f - p[a]) / d[a], tb = (mx[a] - p[a]) / d[a];
                    a0 = fmaxf(a0, fminf(ta, tb)); a1 = fminf(a1, fmaxf(ta, tb));
                } else if (p[a] < -0.5f || p[a] > mx[a]) miss = true;
            }
            if (!miss && a0 < a1) { t0 = fminf(t0, a0); t1 = fmaxf(t1, a1); }
        }
    };
    enum { EMPTY = 0, TRACK = 1, RAYLEIGH = 2, COMPTON = 3 };
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long next_id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;  // the ids this lane starts
    // the photon a lane works on during a visit
    Philox st;
    st.k0 = (unsigned)P.seed; st.k1 = (unsigned)(P.seed >> 32);
    float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1, E = 0, wgt = 0, t = 0, t1 = 0, pmax = 0;
    int n_scat = 0, n_step = 0, n_try = 0, mat = 0;
    auto load = [&](int s) {
        x = W.f[0][s]; y = W.f[1][s]; z = W.f[2][s]; dx = W.f[3][s]; dy = W.f[4][s]; dz = W.f[5][s];
        E = W.f[6][s]; wgt = W.f[7][s]; t = W.f[8][s]; t1 = W.f[9][s]; pmax = W.f[10][s];
        st.c0 = W.u[0][s]; st.c2 = W.u[1][s]; st.c3 = W.u[2][s]; st.o0 = W.u[3][s]; st.o1 = W.u[4][s]; st.o2 = W.u[5][s]; st.o3 = W.u[6][s];
        const unsigned pk = W.u[7][s];
        st.pos = (int)(pk & 0xFFu); n_try = (int)((pk >> 8) & 0xFFu); mat = (int)(pk >> 16);
        n_step = (int)W.u[8][s]; n_scat = (int)W.u[9][s];
    };
    auto store = [&](int s) {
        W.f[0][s] = x; W.f[1][s] = y; W.f[2][s] = z; W.f[3][s] = dx; W.f[4][s] = dy; W.f[5][s] = dz;
        W.f[6][s] = E; W.f[7][s] = wgt; W.f[8][s] = t; W.f[9][s] = t1; W.f[10][s] = pmax;
        W.u[0][s] = st.c0; W.u[1][s] = st.c2; W.u[2][s] = st.c3; W.u[3][s] = st.o0; W.u[4][s] = st.o1; W.u[5][s] = st.o2; W.u[6][s] = st.o3;
        W.u[7][s] = (unsigned)st.pos | ((unsigned)n_try << 8) | ((unsigned)mat << 16);
        W.u[8][s] = (unsigned)n_step; W.u[9][s] = (unsigned)n_scat;
    };
    // the photon left the volumes: tally it if it scattered
    auto leave = [&]() {
        if (n_scat == 0) { c_loc[3] += (double)E * wgt; return; }
        // detector: the plane through the image.  Pixel of a world point X: (uw, vw, w) = P_idx (X, 1); the plane is where the
        // primary rays end, i.e. at depth w = 1 along the principal axis (P_idx is scaled that way, see capi).
        float w0 = P.p_idx[8] * x + P.p_idx[9] * y + P.p_idx[10] * z + P.p_idx[11];
        float wd = P.p_idx[8] * dx + P.p_idx[9] * dy + P.p_idx[10] * dz;
        bool hit = false;
        if (wd > 1e-9f) {
            float s = (1.0f - w0) / wd;  // distance to the detector plane (w == 1)
            if (s > 0.0f) {
                float X = x + s * dx, Y = y + s * dy, Z = z + s * dz;
                float uu = P.p_idx[0] * X + P.p_idx[1] * Y + P.p_idx[2] * Z + P.p_idx[3];
                float vv = P.p_idx[4] * X + P.p_idx[5] * Y + P.p_idx[6] * Z + P.p_idx[7];
                int iu = (int)floorf(uu), iv = (int)floorf(vv);
                if (iu >= 0 && iu < P.W && iv >= 0 && iv < P.H) {
                    hit = true;
                    atomicAdd(P.tally + (size_t)iv * P.W + iu, (unsigned long long)((double)E * (double)wgt * 65536.0 + 0.5));
                }
            }
        }
        c_loc[hit ? 4 : 5] += (double)E * wgt;
    };
    // after a Rayleigh / Compton sample with cos(theta) = cost: energy cut-off, new direction, new span.  Returns the record's tag.
    auto scattered = [&](float cost) -> int {
        if (E < T.e0) { c_loc[2] += (double)E * wgt; return EMPTY; }
        rotate_dir(dx, dy, dz, cost, 6.283185307f * philox_uniform(st));
        n_scat++;
        float t0;
        span(x, y, z, dx, dy, dz, t0, t1);
        t = 0.0f;
        return TRACK;
    };
    // slots whose tag is `what`, in slot order: the first 32 go to W.list; returns how many there are
    auto gather = [&](const int (&tags)[SC_SLOTS / 32], int what) -> int {
        int base = 0;
#pragma unroll
        for (int h = 0; h < SC_SLOTS / 32; h++) {
            const unsigned m = __ballot_sync(0xffffffffu, tags[h] == what);
            const int r = base + __popc(m & lt);
            if (tags[h] == what && r < 32) W.list[r] = (unsigned char)(lane + 32 * h);
            base += __popc(m);
        }
        __syncwarp();
        return base;
    };

#pragma unroll
    for (int h = 0; h < SC_SLOTS / 32; h++) W.tag[lane + 32 * h] = EMPTY;
    __syncwarp();
    for (;;) {
        // ---- census of the pool ---------------------------------------------------------------------------------------------
        int tags[SC_SLOTS / 32];
        int n_empty = 0, n_track = 0, n_ray = 0, n_comp = 0;
#pragma unroll
        for (int h = 0; h < SC_SLOTS / 32; h++) {
            tags[h] = W.tag[lane + 32 * h];
            n_empty += __popc(__ballot_sync(0xffffffffu, tags[h] == EMPTY));
            n_track += __popc(__ballot_sync(0xffffffffu, tags[h] == TRACK));
            n_ray += __popc(__ballot_sync(0xffffffffu, tags[h] == RAYLEIGH));
            n_comp += __popc(__ballot_sync(0xffffffffu, tags[h] == COMPTON));
        }
        const bool more = __any_sync(0xffffffffu, next_id < P.n_photons);
        if (more && n_empty >= 32) {
            // ---- 32 new photons ---------------------------------------------------------------------------------------------
            gather(tags, EMPTY);
            bool placed = false;
            if (next_id < P.n_photons) {
                philox_start(st, P.seed, P.photon_offset + next_id);  // one Philox subsequence per photon: any split is reproducible
                next_id += stride;
                float xi = philox_uniform(st);
                int lo = 0, hi = P.n_bins - 1;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (P.spec_cdf[mid] < xi) lo = mid + 1; else hi = mid; }
                E = P.spec_e_keV[lo] * 1000.0f;
                float u = philox_uniform(st) * P.W, v = philox_uniform(st) * P.H;
                dx = u * P.w2i[0] + v * P.w2i[1] + P.w2i[2]; dy = u * P.w2i[3] + v * P.w2i[4] + P.w2i[5]; dz = u * P.w2i[6] + v * P.w2i[7] + P.w2i[8];
                float rl = sqrtf(dx * dx + dy * dy + dz * dz);
                dx /= rl; dy /= rl; dz /= rl;
                // principal-axis direction has |r| minimal; cos(theta) = r_min / |r|: the weight only needs to be proportional to cos^3
                wgt = 1.0f / (rl * rl * rl);
                x = P.src[0]; y = P.src[1]; z = P.src[2];
                c_loc[0] += (double)E * wgt;
                float t0;
                span(x, y, z, dx, dy, dz, t0, t1);
                if (!(t0 < t1)) c_loc[1] += (double)E * wgt;  // misses every volume
                else { t = t0 + 1e-4f; n_scat = 0; n_step = 0; n_try = 0; mat = 0; pmax = 0.0f; placed = true; }
            }
            const unsigned pm = __ballot_sync(0xffffffffu, placed);
            if (placed) {
                const int s = W.list[__popc(pm & lt)];
                store(s);
                W.tag[s] = TRACK;
            }
            __syncwarp();
            continue;
        }
        if (n_track + n_ray + n_comp == 0) break;  // (more photons with a full pool of empties was handled above)
        // ---- one visit: the tag with the most records --------------------------------------------------------------------------
        const int what = (n_track >= n_comp && n_track >= n_ray) ? TRACK : (n_comp >= n_ray ? COMPTON : RAYLEIGH);
        const int n = min(32, gather(tags, what));
        const bool active = lane < n;
        const int slot = active ? W.list[lane] : 0;
        int state = EMPTY;
        if (active) { load(slot); state = what; }
        if (what == TRACK) {
            // Woodcock tracking
            for (int k = 0; k < SC_STEPS; k++) {
                if (active && state == TRACK && n_step >= 100000) { leave(); state = EMPTY; }  // guard of the one-loop form
                if (active && state == TRACK) {
                    float f = (E - T.e0) / T.de;
                    int ie = max(0, min((int)f, T.n_e - 2));
                    float wq = fminf(fmaxf(f - (float)ie, 0.0f), 1.0f);
                    float smax = T.majorant[ie] + wq * (T.majorant[ie + 1] - T.majorant[ie]);
                    smax *= 1.0001f;
                    t += -__logf(philox_uniform(st)) / smax;
                    n_step++;
                    if (t > t1) { leave(); state = EMPTY; }  // behind the last volume
                    else {
                        const float X = x + t * dx, Y = y + t * dy, Z = z + t * dz;
                        // the volume this point belongs to: smallest priority value among the volumes that contain it
                        int best = -1, best_pr = 0x7fffffff;
                        size_t o = 0;
                        for (int vv = 0; vv < P.V; vv++) {  // (V == 1: the loop runs once; t <= t1 already says "inside" up to rounding)
                            if (!P.enabled[vv] || P.priority[vv] >= best_pr) continue;
                            const float* A = P.ijk[vv];
                            const float qi = A[0] * X + A[1] * Y + A[2] * Z + A[3], qj = A[4] * X + A[5] * Y + A[6] * Z + A[7],
                                        qk = A[8] * X + A[9] * Y + A[10] * Z + A[11];
                            const VolDev& vol = P.vol[vv];
                            if (qi < -0.5f || qi > (float)vol.ni - 0.5f || qj < -0.5f || qj > (float)vol.nj - 0.5f || qk < -0.5f || qk > (float)vol.nk - 0.5f) continue;
                            const int vi = min(max((int)floorf(qi + 0.5f), 0), vol.ni - 1), vj = min(max((int)floorf(qj + 0.5f), 0), vol.nj - 1),
                                      vk = min(max((int)floorf(qk + 0.5f), 0), vol.nk - 1);
                            best = vv; best_pr = P.priority[vv];
                            o = ((size_t)vk * vol.nj + vj) * vol.ni + vi;
                        }
                        if (best >= 0) {  // (between the volumes is vacuum: every interaction there is virtual)
                            // the two voxel reads (a random place in the volume: the long wait of a step) are issued first and
                            // the draw that decides real / virtual is taken while they are in flight
                            const unsigned lb = __ldg(P.vol[best].lab + o);
                            const float rho = __ldg(P.vol[best].dens + o);
                            const float u_virtual = philox_uniform(st);
                            mat = T.mat_of_label[lb];
                            // total inverse mean free path first (five out of six steps are virtual and need nothing else)
                            const float* ma = T.mfp + ((size_t)mat * T.n_e + ie) * 5;
                            const float itot = 1.0f / (ma[3] + wq * (ma[8] - ma[3]));
                            float scale = rho * T.inv_rho_nom[mat];
                            if (!(u_virtual * smax >= itot * scale)) {  // a real interaction
                                const float iray = 1.0f / (ma[0] + wq * (ma[5] - ma[0])), ico = 1.0f / (ma[1] + wq * (ma[6] - ma[1]));
                                pmax = ma[4] + wq * (ma[9] - ma[4]);
                                float r = philox_uniform(st) * itot;
                                x = X; y = Y; z = Z;  // move the photon to the interaction point
                                if (r < iray) state = RAYLEIGH;
                                else if (r < iray + ico) { state = COMPTON; n_try = 0; }
                                else { c_loc[2] += (double)E * wgt; state = EMPTY; }  // photoabsorption
                            }
                        }
                    }
                }
                if (__ballot_sync(0xffffffffu, active && state == TRACK) == 0) break;
            }
        } else if (what == COMPTON) {
            if (active) {
                ComptonTry c;
                if (compton_try(T, mat, E, st, c) || ++n_try == 200) {
                    const float E0 = E;
                    const float cost = compton_finish(T, mat, E, st, c);
                    c_loc[2] += (double)(E0 - E) * wgt; c_loc[7] += 1.0;
                    state = scattered(cost);
                }
            }
        } else {
            if (active) {
                const float cost = sample_rayleigh(T, mat, E, pmax, st);
                c_loc[6] += 1.0;
                state = scattered(cost);
            }
        }
        if (active) {
            if (state != EMPTY) store(slot);
            W.tag[slot] = (unsigned char)state;
        }
        __syncwarp();
    }
    for (int k = 0; k < 8; k++) {
        double vsum = c_loc[k];
        for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if ((threadIdx.x & 31) == 0 && vsum != 0.0) atomicAdd(P.counters + k, vsum);
    }
}

cudaError_t drr_launch_scatter(const ScatterParams& P, int n_sm, cudaStream_t s) {
    // persistent: exactly as many blocks as are resident at once (a lane walks photon ids with the stride of the grid)
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scatter_kernel, 32 * SC_WARPS, 0) != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = SC_MIN_BLOCKS;
    }
    scatter_kernel<<<n_sm * blocks_per_sm, 32 * SC_WARPS, 0, s>>>(P);
    return cudaGetLastError();
}
