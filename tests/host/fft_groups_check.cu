// Host-side check of the group arithmetic of fft_frame.cuh (no GPU needed): for every (avg, pos0, nframes) the groups of a
// call must tile the call's frames exactly, never straddle a canonical group of the averaging block, start from the carry
// row exactly when the call begins inside a canonical group, and be complete exactly when they reach its end.
#include <stdio.h>
#include <stdlib.h>

#include "fft_frame.cuh"

int main() {
    const int avgs[] = {1, 2, 7, 12, 13, 25, 64, 100, 257};
    long long checked = 0;
    unsigned seed = 12345u;
    for (int avg : avgs) {
        const int ng = (avg + 11) / 12;              // fft_config
        const int G = (avg + ng - 1) / ng;
        const int gpb = (avg + G - 1) / G;
        for (int pos0 = 0; pos0 < avg; ++pos0) {
            for (int trial = 0; trial < 40; ++trial) {
                seed = seed * 1664525u + 1013904223u;
                const int nframes = 1 + (int)((seed >> 8) % (unsigned)(5 * avg + 3));
                const int n = rcb::fft_frame_ngroups(nframes, pos0, avg, G, gpb);
                int expect_f0 = 0;
                for (int gi = 0; gi < n; ++gi) {
                    int f0, f1;
                    bool fc, complete;
                    rcb::fft_frame_group(gi, nframes, pos0, avg, G, gpb, &f0, &f1, &fc, &complete);
                    const long long a = (long long)pos0 + f0, b = (long long)pos0 + f1 - 1;  // block-0-relative, inclusive
                    bool ok = (f0 == expect_f0) && (f0 < f1) && (f1 <= nframes);
                    ok = ok && (a / avg == b / avg) && ((a % avg) / G == (b % avg) / G);
                    ok = ok && (fc == (gi == 0 && (pos0 % G) != 0));
                    const long long e = (long long)pos0 + f1;                                  // one past the group's last frame
                    const bool at_end = (e % avg == 0) || ((e % avg) % G == 0);
                    ok = ok && (complete == at_end) && (complete || gi == n - 1);
                    if (!ok) {
                        printf("FAIL avg %d G %d pos0 %d nframes %d gi %d of %d: [%d, %d) carry %d complete %d\n", avg, G, pos0,
                               nframes, gi, n, f0, f1, (int)fc, (int)complete);
                        return 1;
                    }
                    expect_f0 = f1;
                    ++checked;
                }
                if (expect_f0 != nframes) {
                    printf("FAIL avg %d pos0 %d nframes %d: groups end at %d\n", avg, pos0, nframes, expect_f0);
                    return 1;
                }
                // the group after the last one must be empty
                int f0, f1;
                bool fc, complete;
                rcb::fft_frame_group(n, nframes, pos0, avg, G, gpb, &f0, &f1, &fc, &complete);
                if (f0 < f1) {
                    printf("FAIL avg %d pos0 %d nframes %d: group %d is not empty\n", avg, pos0, nframes, n);
                    return 1;
                }
            }
        }
    }
    printf("ok %lld groups\n", checked);
    return 0;
}
