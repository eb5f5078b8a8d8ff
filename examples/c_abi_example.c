/* Minimal C caller of the drop-in boundary (include/metdet_b200.h): what a non-Python host would link against.
 * Build:  gcc -std=c99 -Iinclude examples/c_abi_example.c -Lmetdetpy_b200 -lmetdet_b200 -Wl,-rpath,$PWD/metdetpy_b200 -o c_abi_example
 * Runs T x { update(frame); detect() } through mdb_detect_batch on synthetic frames (a bright bar moving over noise)
 * and prints the frames that produced lines.  Without a CUDA device it reports that and exits 0: there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "metdet_b200.h"

int main(void) {
    enum { W = 320, H = 240, N = 5, T = 24 };
    printf("libmetdet_b200 ABI version %d, %d CUDA device(s)\n", mdb_version(), mdb_device_count());
    if (mdb_device_count() < 1) {
        printf("no CUDA device: nothing to run (the library has no CPU fallback)\n");
        return 0;
    }
    mdb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.width = W; cfg.height = H; cfg.window = N;
    cfg.adaptive = 1; cfg.init_value = 7; cfg.sensitivity = MDB_SENS_NORMAL; cfg.nz_interval = 2;
    cfg.roi[0] = H / 3; cfg.roi[1] = W / 3; cfg.roi[2] = 2 * H / 3; cfg.roi[3] = 2 * W / 3; /* SNR_SW.std_roi (r0,c0,r1,c1) */
    cfg.hough_threshold = 10; cfg.hough_min_len = 10; cfg.hough_max_gap = 10;
    cfg.dy_mask = 1; cfg.max_batch = T; cfg.device = 0;
    uint8_t *mask = malloc((size_t)W * H), *frames = malloc((size_t)T * W * H);
    memset(mask, 1, (size_t)W * H);
    unsigned s = 12345u;
    for (int t = 0; t < T; t++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                s = s * 1664525u + 1013904223u;
                int v = 30 + (int)((s >> 24) & 3);
                if (t >= 8 && t < 16 && y >= 100 && y < 102 && x >= 40 + 12 * (t - 8) && x < 64 + 12 * (t - 8)) v = 200;
                frames[((size_t)t * H + y) * W + x] = (uint8_t)v;
            }
    mdb_handle h = NULL;
    if (mdb_create(&cfg, mask, &h) != MDB_OK) { fprintf(stderr, "mdb_create: %s\n", mdb_last_error()); return 1; }
    mdb_frame_info *infos = calloc(T, sizeof *infos);
    int32_t *lines = calloc((size_t)T * MDB_MAX_LINES * 4, sizeof *lines);
    double *prob = calloc((size_t)T * MDB_MAX_LINES, sizeof *prob);
    if (mdb_detect_batch(h, frames, T, 0, infos, lines, prob, NULL, NULL, 0) != MDB_OK) {
        fprintf(stderr, "mdb_detect_batch: %s\n", mdb_last_error());
        return 1;
    }
    for (int t = 0; t < T; t++)
        if (infos[t].n_lines > 0) {
            const int32_t *l = lines + (size_t)t * MDB_MAX_LINES * 4;
            printf("frame %2d: threshold %d, %d on-pixels, %d raw segments, %d lines, first (%d,%d)-(%d,%d) nonline_prob %.3f\n", t,
                   infos[t].bi_threshold, infos[t].n_on, infos[t].lines_num, infos[t].n_lines, l[0], l[1], l[2], l[3],
                   prob[(size_t)t * MDB_MAX_LINES]);
        }
    mdb_destroy(h);
    free(mask); free(frames); free(infos); free(lines); free(prob);
    return 0;
}
