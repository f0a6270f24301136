/* Plain-C oracle for the fused Dice+BCE loss partial sums and the Dice / IoU integer counters.
 *
 * TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Never linked into the product library.
 *
 * Restates (formulas in oracle/loss_metrics.py's header, with the call sites):
 *   monai.losses.DiceCELoss(sigmoid=True, lambda_dice=1, lambda_ce=0.2)
 *       /root/reference/configs/model/maple_clipseg.yaml:29-33
 *   torchmetrics.Dice(threshold, zero_division=1, average="samples")  ->  per-sample tp/fp/fn with p >= thr
 *   torchmetrics.JaccardIndex(task="binary", threshold)              ->  global [[tn,fp],[fn,tp]] with p >  thr
 *       /root/reference/src/models/image_text_mask_module.py:284-298, targets = mask.long() (:107)
 *
 * p = fl32(1 / fl32(1 + fl32(exp(-x)))) with exp evaluated in double and rounded once (a correctly
 * rounded fp32 exp).  Sums are accumulated in double; the CUDA kernel accumulates fp32 per thread and
 * fp64 across threads, so the float outputs are compared with a tolerance and the integers bit-exactly.
 */
#include <math.h>
#include <stdint.h>

static float sigmoid_f32(float x)
{
    float e = (float)exp(-(double)x);
    float s = 1.0f + e;
    return 1.0f / s;
}

/* parts[b] = {I, P, G, bce_sum}; counts[b] = {tp, fp, fn} (>=); conf = {tn, fp, fn, tp} (>) summed over b */
void oracle_dicebce_metrics(const float *logits, const float *mask, long long B, long long N, float thr,
                            double *parts, int64_t *counts, int64_t *conf)
{
    conf[0] = conf[1] = conf[2] = conf[3] = 0;
    for (long long b = 0; b < B; ++b) {
        double I = 0, P = 0, G = 0, bce = 0;
        int64_t tp = 0, fp = 0, fn = 0;
        for (long long i = 0; i < N; ++i) {
            float x = logits[b * N + i], y = mask[b * N + i];
            float p = sigmoid_f32(x);
            I += (double)p * y;
            P += p;
            G += y;
            /* BCEWithLogits: max(x,0) - x*y + log(1 + exp(-|x|)) */
            bce += (x > 0 ? (double)x : 0.0) - (double)x * y + log1p(exp(-fabs((double)x)));
            int t = (int)(long long)y; /* mask.long(): truncation */
            int ge = p >= thr, gt = p > thr;
            tp += ge & t;
            fp += ge & !t;
            fn += (!ge) & t;
            conf[t * 2 + gt] += 1;
        }
        parts[b * 4 + 0] = I;
        parts[b * 4 + 1] = P;
        parts[b * 4 + 2] = G;
        parts[b * 4 + 3] = bce;
        counts[b * 3 + 0] = tp;
        counts[b * 3 + 1] = fp;
        counts[b * 3 + 2] = fn;
    }
}
