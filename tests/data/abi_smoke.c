/* Plain-C client of the C ABI (no GPU needed): the header must be valid C, the host-side entry points must work
 * and the device entry points must fail loudly.  Built and run by tests/test_host_cpu.py. */
#include "b200call.h"

#include <stdio.h>
#include <string.h>

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            fprintf(stderr, "abi_smoke: %s failed (line %d): %s\n", #cond, __LINE__, b200_last_error()); \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main(void) {
    CHECK(b200_version() != NULL && strlen(b200_version()) > 0);

    /* chunk offsets: tests/ChunkTest.cpp golden vector */
    uint64_t offs[8], n = 0;
    CHECK(b200_generate_chunks(3 * 9996, 9996, 6, 498, offs, 8, &n) == B200_OK);
    CHECK(n == 4 && offs[0] == 0 && offs[1] == 9498 && offs[2] == 18996 && offs[3] == 19992);
    CHECK(b200_generate_chunks(0, 9996, 6, 498, offs, 8, &n) == B200_ERR_INVALID);
    uint64_t iv[8];
    CHECK(b200_generate_variable_chunks(9996 + 1, 9996, 6, 498, iv, 4, &n) == B200_OK);
    CHECK(n == 2 && iv[0] == 0 && iv[1] == 5244 && iv[2] == 4752 && iv[3] == 9997);

    /* stitching two chunks of 10 blocks overlapping by 4 blocks (stride 1) */
    const uint8_t m0[10] = {1, 0, 1, 0, 1, 0, 1, 0, 1, 0}, m1[10] = {1, 1, 0, 0, 1, 0, 0, 1, 0, 1};
    b200_called_chunk c[2];
    memset(c, 0, sizeof(c));
    c[0].input_offset = 0;  c[0].raw_chunk_size = 10; c[0].moves = m0; c[0].n_moves = 10;
    c[0].sequence = "ACGTA"; c[0].qstring = "!!!!!"; c[0].n_bases = 5;
    c[1].input_offset = 6;  c[1].raw_chunk_size = 10; c[1].moves = m1; c[1].n_moves = 10;
    c[1].sequence = "CCGTA"; c[1].qstring = "#####"; c[1].n_bases = 5;
    uint8_t mo[20];
    char so[11] = {0}, qo[11] = {0};
    uint64_t nm = 0, nb = 0;
    CHECK(b200_stitch_chunks(c, 2, 16, 1, mo, so, qo, &nm, &nb) == B200_OK);
    /* chunk 0 keeps blocks 0..7 (bases A,C,G,T), chunk 1 keeps blocks 2..9 (bases G,T,A) */
    CHECK(nm == 16 && nb == 7 && strncmp(so, "ACGTGTA", 7) == 0 && strncmp(qo, "!!!!###", 7) == 0);

    /* batch-size selection rule */
    const int32_t bs[4] = {64, 128, 192, 256};
    const float ms[4] = {1.0f, 0.6f, 0.65f, 0.5f};
    int32_t sel = 0;
    CHECK(b200_select_batch_size(bs, ms, 4, 10240, 64, 0.0f, &sel) == B200_OK && sel == 256);
    CHECK(b200_select_batch_size(bs, ms, 4, 200, 64, 0.0f, &sel) == B200_OK && sel == 128);

    /* device entry points: no CPU fallback */
    if (b200_device_count() == 0) {
        b200_model_desc desc;
        memset(&desc, 0, sizeof(desc));
        b200_engine* e = NULL;
        CHECK(b200_engine_create(&desc, NULL, 0, 0, &e) != B200_OK && e == NULL);
        CHECK(strlen(b200_last_error()) > 0);
    }
    printf("abi_smoke ok\n");
    return 0;
}
