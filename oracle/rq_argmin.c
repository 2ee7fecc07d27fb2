/* Oracle (TEST INFRASTRUCTURE): scalar C restatement of the RQ-VAE residual nearest-codebook search.
 * Follows /root/reference/genrec/models/rqvae.py:185-199 (dist = |x|^2 + |c|^2 - 2 x.c, first-min argmin)
 * and :397-405 (res <- res - codebook[id], level after level).  fp32 arithmetic, no FMA contraction, so it is
 * the "plain reading" of the formula; exact ties / near-ties are compared against torch in the tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void oracle_rq_residual_argmin(const float* res_in, const float* codebooks, int64_t n, int d, int k, int levels,
                               int64_t* ids, float* res_out) {
    float* r = (float*)malloc(sizeof(float) * (size_t)d);
    float* cn = (float*)malloc(sizeof(float) * (size_t)k * (size_t)levels);
    for (int l = 0; l < levels; ++l)
        for (int c = 0; c < k; ++c) {
            const float* cv = codebooks + ((size_t)l * k + c) * d;
            float s = 0.f;
            for (int j = 0; j < d; ++j) s += cv[j] * cv[j];
            cn[(size_t)l * k + c] = s;
        }
    for (int64_t i = 0; i < n; ++i) {
        memcpy(r, res_in + (size_t)i * d, sizeof(float) * (size_t)d);
        for (int l = 0; l < levels; ++l) {
            float xn = 0.f;
            for (int j = 0; j < d; ++j) xn += r[j] * r[j];
            int best = 0;
            float bestd = 0.f;
            for (int c = 0; c < k; ++c) {
                const float* cv = codebooks + ((size_t)l * k + c) * d;
                float dot = 0.f;
                for (int j = 0; j < d; ++j) dot += r[j] * cv[j];
                float dist = (xn + cn[(size_t)l * k + c]) - 2.f * dot;
                if (c == 0 || dist < bestd) { bestd = dist; best = c; }
            }
            ids[(size_t)i * levels + l] = best;
            const float* cv = codebooks + ((size_t)l * k + best) * d;
            for (int j = 0; j < d; ++j) r[j] -= cv[j];
        }
        memcpy(res_out + (size_t)i * d, r, sizeof(float) * (size_t)d);
    }
    free(r);
    free(cn);
}
