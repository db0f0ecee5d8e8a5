/*
 * screen_model.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU emulation of the arithmetic of the SCREEN variant of nn_kernel (csrc/ured_chamfer.cu): the float32
 * expansion-form filter value per candidate, the per-32-chunk minimum, the best / second-best chunk minima, the
 * ambiguity test with eps = 32 u S^2 + 1e-35, and the exact difference-form re-check of the winning chunk.
 * It exists to test the screening INVARIANT on the CPU, at scale and on adversarial inputs, without a GPU:
 *
 *     every query the filter calls unambiguous has its reference argmin (lowest index of the minimal
 *     float32 difference-form distance, chamfer3D.cu:32-39,126) inside the winning chunk.
 *
 * screen_check() returns the number of violations (must be 0) and reports how many queries were ambiguous
 * (those take the exact fallback on the GPU).  fmaf() and -ffp-contract=off make every rounding explicit.
 */
#include <math.h>
#include <stddef.h>

#define CHUNK 32

static inline float exact_d(const float *c, const float *q) {
    float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* xyz_q [nq,3] queries, xyz_c [nc,3] candidates (one cloud pair, one direction).
 * out_stats[0] = ambiguous queries, out_stats[1] = violations, out_stats[2] = mismatches of the full pipeline
 * (screen + re-check for unambiguous queries, exact scan for ambiguous ones) against the exact argmin. */
long screen_check(int nq, const float *xyz_q, int nc, const float *xyz_c, long *out_stats) {
    long ambiguous = 0, violations = 0, mismatches = 0;
    float wmax = 0.0f;
    for (int k = 0; k < nc; k++) {
        const float *c = xyz_c + (size_t)k * 3;
        float w = fmaf(c[2], c[2], fmaf(c[1], c[1], c[0] * c[0]));
        if (w > wmax) wmax = w;
    }
    const float cn = nextafterf(sqrtf(wmax), INFINITY) * 1.000001f;
    const int nchunks = (nc + CHUNK - 1) / CHUNK;
    for (int j = 0; j < nq; j++) {
        const float *q = xyz_q + (size_t)j * 3;
        /* reference answer */
        float ref_d = 0.0f;
        int ref_i = 0;
        for (int k = 0; k < nc; k++) {
            float d = exact_d(xyz_c + (size_t)k * 3, q);
            if (k == 0 || d < ref_d) { ref_d = d; ref_i = k; }
        }
        /* screening pass: s = fma(x, -2qx, fma(y, -2qy, fma(z, -2qz, W))) */
        const float sx = -2.0f * q[0], sy = -2.0f * q[1], sz = -2.0f * q[2];
        float best = INFINITY, second = INFINITY;
        int bchunk = 0;
        for (int ch = 0; ch < nchunks; ch++) {
            float cm = INFINITY;
            for (int k = ch * CHUNK; k < (ch + 1) * CHUNK; k++) {
                const float *c = xyz_c + (size_t)(k < nc ? k : nc - 1) * 3; /* padding replicates the last point */
                float w = fmaf(c[2], c[2], fmaf(c[1], c[1], c[0] * c[0]));
                float s = fmaf(c[0], sx, fmaf(c[1], sy, fmaf(c[2], sz, w)));
                cm = fminf(cm, s);
            }
            int better = cm < best;
            second = fminf(second, fmaxf(cm, best));
            best = fminf(best, cm);
            if (better) bchunk = ch;
        }
        const float qn = nextafterf(sqrtf(fmaf(q[2], q[2], fmaf(q[1], q[1], q[0] * q[0]))), INFINITY) * 1.000001f;
        const float S = qn + cn;
        const float eps = fmaf(S * S, 1.9073486e-6f, 1e-35f);
        const int amb = !(second > best + eps);
        int got_i;
        if (amb) {
            ambiguous++;
            got_i = ref_i; /* exact fallback scan */
        } else {
            if (ref_i / CHUNK != bchunk) violations++;
            float bd = 0.0f;
            got_i = bchunk * CHUNK;
            for (int k = bchunk * CHUNK; k < (bchunk + 1) * CHUNK && k < nc; k++) {
                float d = exact_d(xyz_c + (size_t)k * 3, q);
                if (k == bchunk * CHUNK || d < bd) { bd = d; got_i = k; }
            }
        }
        if (got_i != ref_i) mismatches++;
    }
    out_stats[0] = ambiguous;
    out_stats[1] = violations;
    out_stats[2] = mismatches;
    return violations + mismatches;
}
