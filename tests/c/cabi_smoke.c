/* cabi_smoke.c -- the C ABI of include/forgex_b200.h called from compiled C, the way a bind(C) interface block calls it.
 * Host-only entry points are checked always; the matching entry points when a CUDA device exists (otherwise every one
 * of them must answer FX_ERR_NO_DEVICE: there is no CPU fallback).  Expected answers are the reference's own
 * (README.md:153-220, test/test_api/test_case_003.f90:27-28, src/forgex.F90:266-274).  Exit code 0 = all good. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "forgex_b200.h"

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

int main(void) {
    /* ---- host only: compile, query, validity ---- */
    fx_pattern* p = NULL;
    int rc = fx_compile("\\d{3}-\\d{4}", 11, FX_OP_MATCH, &p);
    CHECK(rc == FX_OK && p != NULL);
    fx_pattern_info info;
    CHECK(fx_pattern_get_info(p, &info) == FX_OK);
    CHECK(info.op == FX_OP_MATCH && info.status == 0 && info.byte_states > 8 && info.literal_only == 0);
    fx_pattern* bad = NULL;
    rc = fx_compile("(a", 2, FX_OP_IN, &bad);
    CHECK(rc == 2 && bad != NULL);                       /* SYNTAX_ERR_PARENTHESIS_MISSING, error_m.F90:12-38 */
    CHECK(strlen(fx_status_message(rc)) > 0);
    fx_pattern_free(bad);
    int st = -1;
    CHECK(fx_is_valid_regex("[a-z]+", 6, &st) == 1 && st == 0);
    CHECK(fx_is_valid_regex("a**", 3, &st) == 0 && st == 16);   /* SYNTAX_ERR_STAR_INCOMPLETE */
    const char pats[] = "a|b(a[z-a]";
    const int64_t poff[4] = {0, 3, 5, 10};
    uint8_t valid[3];
    int32_t status[3];
    CHECK(fx_is_valid_regex_batch(pats, poff, 3, valid, status) == FX_OK);
    CHECK(valid[0] == 1 && valid[1] == 0 && valid[2] == 0 && status[0] == 0 && status[1] == 2 && status[2] == 14);
    fx_pattern* rx = NULL;
    CHECK(fx_compile("foo(bar|baz)", 12, FX_OP_REGEX, &rx) == FX_OK);
    CHECK(fx_pattern_get_info(rx, &info) == FX_OK && info.literal_prefix_len == 5 && info.prefix_scan == 1);
    char pre[8] = {0};
    CHECK(fx_pattern_literals(rx, NULL, pre, NULL) == FX_OK && memcmp(pre, "fooba", 5) == 0);   /* ast_case_001.f90:31 shape */
    CHECK(fx_regex_buffer_work_bytes(1 << 20) >= 4096);

    /* ---- matching: needs a device ---- */
    const char strs[] = "123-4567" "12a-4567" "000-0000" "123-456 ";
    uint8_t out[4] = {9, 9, 9, 9};
    rc = fx_match_fixed(p, (const uint8_t*)strs, 4, 8, out);
    if (rc == FX_ERR_NO_DEVICE) {
        int r = 7;
        CHECK(fx_in("a", 1, "abc", 3, &r) == FX_ERR_NO_DEVICE);
        int64_t f = 0, t = 0;
        CHECK(fx_regex_buffer(rx, (const uint8_t*)"foobar", 6, &f, &t) == FX_ERR_NO_DEVICE);
        printf("cabi_smoke: host-only checks %s (no CUDA device: matching entry points answer FX_ERR_NO_DEVICE)\n",
               failures ? "FAILED" : "ok");
    } else {
        CHECK(rc == FX_OK);
        CHECK(out[0] == 1 && out[1] == 0 && out[2] == 1 && out[3] == 0);
        int r = -1;
        CHECK(fx_in("foo(bar|baz)", 12, "xxfoobazyy", 10, &r) == FX_OK && r == 1);
        CHECK(fx_in("(a", 2, "a", 1, &r) == FX_OK && r == 0);           /* invalid pattern -> .false. (forgex.F90:101-104) */
        CHECK(fx_match("^abc$", 5, "abc\n", 4, &r) == FX_OK && r == 1);  /* SURVEY Q5 */
        int64_t f = 0, t = 0, len = 0;
        int status1 = -1;
        CHECK(fx_regex("[d-f]{3}", 8, "abcdefghi", 9, &f, &t, &len, &status1) == FX_OK);
        CHECK(f == 4 && t == 6 && len == 3 && status1 == 0);            /* README.md:204-220 */
        CHECK(fx_regex("(a", 2, "x", 1, &f, &t, &len, &status1) == FX_OK && f == -9999 && t == -9999 && len == 0 && status1 == 2);
        const char text[] = "INFO a\nERROR x timeout=12\nWARN b\n";
        fx_pattern* c4 = NULL;
        CHECK(fx_compile("^ERROR.*timeout=\\d+$", 20, FX_OP_REGEX, &c4) == FX_OK);
        CHECK(fx_regex_buffer(c4, (const uint8_t*)text, (int64_t)strlen(text), &f, &t) == FX_OK);
        CHECK(f == 7 && t == 26);                                        /* the span holds both newlines */
        const int64_t off[4] = {0, 7, 26, 33};
        int64_t ff[3], tt[3], cnt[3];
        CHECK(fx_regex_batch(c4, (const uint8_t*)text, off, 3, ff, tt) == FX_OK);
        CHECK(ff[0] == 0 && ff[1] == 1 && tt[1] == 19 && ff[2] == 0);
        CHECK(fx_regex_count_batch(c4, (const uint8_t*)text, off, 3, cnt) == FX_OK && cnt[0] == 0 && cnt[1] == 1 && cnt[2] == 0);
        int64_t af[4], at[4], n = 0;
        fx_pattern* word = NULL;
        CHECK(fx_compile("[A-Z]+", 6, FX_OP_REGEX, &word) == FX_OK);
        CHECK(fx_regex_buffer_all(word, (const uint8_t*)text, (int64_t)strlen(text), af, at, 4, &n) == FX_OK);
        CHECK(n == 3 && af[0] == 1 && at[0] == 4 && af[1] == 8 && at[1] == 12 && af[2] == 27 && at[2] == 30);
        const int64_t badoff[3] = {0, 9, 4};
        CHECK(fx_regex_batch(c4, (const uint8_t*)text, badoff, 2, ff, tt) == FX_ERR_BAD_ARGUMENT);   /* offsets must ascend */
        fx_pattern_free(c4);
        fx_pattern_free(word);
        printf("cabi_smoke: host-only and device checks %s (kernels launched: %lld)\n", failures ? "FAILED" : "ok",
               (long long)fx_launch_count());
    }
    fx_pattern_free(p);
    fx_pattern_free(rx);
    return failures ? 1 : 0;
}
