/* TEST INFRASTRUCTURE ONLY — C restatement of the reference's Vose alias-table builder.
 *
 * Follows /root/reference/code/nce/alias_multinomial.py:40-73 step for step so the tables are bit-identical to
 * what the reference's O(V) Python loop produces (pinned by tests/golden/alias_*.pt):
 *   :48-53  self_prob[i] = K * prob[i]  in float32; i goes to `smaller` if < 1.0 else `larger`  (scan order 0..K-1)
 *   :58-68  pop the LAST element of each list; alias[small] = large;
 *           prob[large] = (prob[large] - 1.0) + prob[small]  (two float32 roundings);
 *           push `large` on the end of smaller/larger
 *   :70-71  every leftover gets prob = 1
 * Build:  gcc -O2 -ffp-contract=off -shared -fPIC alias_vose.c -o libmap_oracle.so   (see oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>

int oracle_alias_build(const float* probs, int64_t K, float* out_prob, int64_t* out_alias) {
    int64_t* smaller = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K > 0 ? K : 1));
    int64_t* larger = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K > 0 ? K : 1));
    if (!smaller || !larger) { free(smaller); free(larger); return -1; }
    int64_t ns = 0, nl = 0;
    const float Kf = (float)K;
    for (int64_t i = 0; i < K; ++i) {
        volatile float p = Kf * probs[i];
        out_prob[i] = p;
        out_alias[i] = 0;
        if (p < 1.0f) smaller[ns++] = i; else larger[nl++] = i;
    }
    while (ns > 0 && nl > 0) {
        int64_t small = smaller[--ns];
        int64_t large = larger[--nl];
        out_alias[small] = large;
        volatile float t = out_prob[large] - 1.0f;
        volatile float q = t + out_prob[small];
        out_prob[large] = q;
        if (q < 1.0f) smaller[ns++] = large; else larger[nl++] = large;
    }
    for (int64_t i = 0; i < ns; ++i) out_prob[smaller[i]] = 1.0f;
    for (int64_t i = 0; i < nl; ++i) out_prob[larger[i]] = 1.0f;
    free(smaller); free(larger);
    return 0;
}
