/* A plain C host of the drop-in boundary (include/snk_b200.h): loads a small synthetic epoch voice, sets stream weights,
 * runs the greedy joint search (the reference's Synthesiser.greedy_joint_search, script/synth_simple.py:458-503) for one
 * utterance made of database frames and checks the known answer the reference itself asserts (synth_simple.py:909-928:
 * database frames in, consecutive units out).  No Python, no torch: the library needs nothing but the CUDA driver.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -Lsnickery_b200/_lib -lsnk_b200 -Wl,-rpath,$PWD/snickery_b200/_lib -lm -o c_host
 *
 * Exit codes: 0 = path as expected, 2 = no B200 visible (the error text of snk_last_error() is printed), 1 = anything else. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include "snk_b200.h"

enum { N = 6000, DT = 61, DJ = 151, M = 6, STEPS = 10, START = 1200 };

static float frand(uint32_t *s) {            /* xorshift: deterministic, no libc dependence on RAND_MAX */
    *s ^= *s << 13; *s ^= *s >> 17; *s ^= *s << 5;
    return (float)(*s & 0xFFFFFF) / (float)0x1000000 - 0.5f;
}

int main(void) {
    if (snk_device_count() < 1) {
        fprintf(stderr, "no CUDA device: %s\n", snk_last_error());
        return 2;
    }
    float *F = malloc(sizeof(float) * N * DT), *Jc = malloc(sizeof(float) * (N + 1) * DJ);
    double wt[DT], wj[DJ], *targets = malloc(sizeof(double) * STEPS * M * DT);
    int64_t lens[1] = {STEPS * M}, start[1] = {START}, path[STEPS];
    double dist[STEPS];
    uint32_t seed = 2463534242u;
    if (!F || !Jc || !targets) return 1;
    /* a smooth random walk per column, like consecutive speech frames; join context u = frame u - 1's full vector */
    for (int c = 0; c < DT; ++c) F[c] = frand(&seed);
    for (int u = 1; u < N; ++u)
        for (int c = 0; c < DT; ++c) F[u * DT + c] = 0.95f * F[(u - 1) * DT + c] + 0.3f * frand(&seed);
    for (int c = 0; c < DJ; ++c) Jc[c] = frand(&seed);
    for (int u = 1; u <= N; ++u)
        for (int c = 0; c < DJ; ++c) Jc[u * DJ + c] = 0.95f * Jc[(u - 1) * DJ + c] + 0.3f * frand(&seed);
    for (int c = 0; c < DT; ++c) wt[c] = 0.5 / sqrt((double)DT);
    for (int c = 0; c < DJ; ++c) wj[c] = 0.2 / sqrt((double)DJ);

    snk_db *db = NULL;
    if (snk_db_create(&db, 0, N, DT, DJ, M, F, Jc, SNK_LAYOUT_SIMPLE)) {
        fprintf(stderr, "snk_db_create: %s\n", snk_last_error());
        return 2;
    }
    if (snk_db_set_weights(db, wt, wj)) { fprintf(stderr, "snk_db_set_weights: %s\n", snk_last_error()); return 1; }
    /* the utterance: weighted database frames START .. START + STEPS * M, as weight(train_unit_features) would give them */
    for (int t = 0; t < STEPS * M; ++t)
        for (int c = 0; c < DT; ++c) targets[t * DT + c] = (double)F[(START + t) * DT + c] * wt[c];
    if (snk_greedy_batch(db, targets, lens, 1, start, path, dist)) {
        fprintf(stderr, "snk_greedy_batch: %s\n", snk_last_error());
        return 1;
    }
    int ok = 1;
    for (int s = 0; s < STEPS; ++s) {
        printf("step %d: unit %lld  distance %.3g\n", s, (long long)path[s], dist[s]);
        ok = ok && path[s] == START + s * M && dist[s] == 0.0;
    }
    snk_db_destroy(db);
    free(F); free(Jc); free(targets);
    printf(ok ? "identity path recovered\n" : "UNEXPECTED PATH\n");
    return ok ? 0 : 1;
}
