/* Host-side check of alpha_jump (deepdrr_b200/csrc/drr_device.cuh): the same integer-on-the-bit-pattern algorithm in plain C
 * against n sequential float additions (what projectKernel does, project_kernel.cu:552).  Built and run by
 * tests/test_alpha_jump.py; exit code 0 = every case bit-identical. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
// alpha after n sequential fp32 additions of step (alpha > 0 normal, step > 0)
static float alpha_jump(float alpha, float step, int n) {
    while (n > 0) {
        volatile float nx = alpha + step;
        float next = nx;
        uint32_t ua = f2u(alpha), un = f2u(next);
        if ((ua >> 23) != (un >> 23) || n == 1) { alpha = next; n--; continue; }   // leaves the binade (or last step): single step
        // same binade: increment in ulps
        uint32_t D = un - ua;                    // integer number of ulps added (both in the same binade)
        if (D == 0) return alpha;                // step below half an ulp: alpha is stuck for good
        volatile float dv = next - alpha; float d = dv;   // exact
        volatile float rv = step - d; float r = rv;       // exact remainder, |r| <= ulp/2
        float ulp = u2f(((ua >> 23) - 23) << 23);
        if (fabsf(r) * 2.0f == ulp) { alpha = next; n--; continue; }               // tie: parity dependent, step singly
        uint32_t A = ua & 0x7FFFFF;              // mantissa offset within the binade, in ulps
        uint32_t room = 0x7FFFFF - A;            // ulps left before the top of the binade
        uint32_t m = room / D;                   // steps that stay inside the binade
        if (m == 0) { alpha = next; n--; continue; }
        if (m > (uint32_t)n) m = n;
        alpha = u2f(ua + m * D);
        n -= m;
    }
    return alpha;
}
static float seq(float alpha, float step, int n){ for(int i=0;i<n;i++){ volatile float t=alpha+step; alpha=t;} return alpha; }
int main(){
    srand(1); long bad=0, tot=0;
    float steps[] = {0.1f, 0.05f, 0.25f, 0.5f, 0.3f, 1.0f, 0.125f, 0.2f, 0.0625f, 0.07f, 0.1234567f, 3.0f, 1e-3f};
    for (int si=0; si<13; si++) for (int it=0; it<20000; it++){
        float a = 0.05f + (rand()/(float)RAND_MAX)*((it%3)?5.0f:900.0f);
        int n = rand()%9000;
        float x = alpha_jump(a, steps[si], n), y = seq(a, steps[si], n);
        tot++; if (f2u(x)!=f2u(y)) { if (bad<10) printf("BAD a=%.9g step=%.9g n=%d jump=%.9g seq=%.9g\n", a, steps[si], n, x, y); bad++; }
    }
    // random steps
    for (int it=0; it<200000; it++){
        float a = 0.01f + (rand()/(float)RAND_MAX)*1200.0f; float st = 1e-3f + (rand()/(float)RAND_MAX)*2.0f; int n = rand()%5000;
        if (it%5==0) st = ldexpf(1.0f, -(rand()%12));   // powers of two: exact / tie prone
        if (it%7==0) st = ldexpf((float)(1+rand()%7), -(rand()%26));
        float x = alpha_jump(a, st, n), y = seq(a, st, n);
        tot++; if (f2u(x)!=f2u(y)) { if (bad<20) printf("BAD a=%.9g step=%.9g n=%d jump=%.9g seq=%.9g\n", a, st, n, x, y); bad++; }
    }
    printf("tot %ld bad %ld\n", tot, bad); return bad!=0;
}
