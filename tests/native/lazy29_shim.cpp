// Host build of thaler_study_b200/csrc/lazy29.hpp for tests/test_lazy29_host.py (the same source the device compiles).
#include "../../thaler_study_b200/csrc/lazy29.hpp"
using namespace scb::l29;
extern "C" {
int l29_make_desc(const uint64_t* p64, uint32_t bits, Desc29* out) { return make_desc(p64, bits, out) ? 1 : 0; }
void l29_from_words(const uint32_t* w, uint32_t* limbs) {
    uint32_t ww[8];
    for (int i = 0; i < 8; ++i) ww[i] = w[i];
    L9 r = from_words(ww);
    for (int j = 0; j < NL; ++j) limbs[j] = r.l[j];
}
void l29_to_words(const uint32_t* limbs, uint32_t* w9) {
    L9 a;
    for (int j = 0; j < NL; ++j) a.l[j] = limbs[j];
    uint32_t w[8], top;
    to_words(a, w, top);
    for (int i = 0; i < 8; ++i) w9[i] = w[i];
    w9[8] = top;
}
void l29_normalise(const uint32_t* in, uint32_t* out) {
    L9 a;
    for (int j = 0; j < NL; ++j) a.l[j] = in[j];
    L9 r = normalise(a);
    for (int j = 0; j < NL; ++j) out[j] = r.l[j];
}
void l29_sub_kp(const Desc29* d, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    L9 x, y;
    for (int j = 0; j < NL; ++j) { x.l[j] = a[j]; y.l[j] = b[j]; }
    L9 r = sub_kp(*d, x, y);
    for (int j = 0; j < NL; ++j) out[j] = r.l[j];
}
// returns the largest column value seen (head-room check) through max_col
void l29_mont(const Desc29* d, const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t* max_col) {
    L9 x, y;
    for (int j = 0; j < NL; ++j) { x.l[j] = a[j]; y.l[j] = b[j]; }
    uint64_t t[NL];
    mont_cols(*d, x, y, t);
    uint64_t m = 0;
    for (int j = 0; j < NL; ++j) m = t[j] > m ? t[j] : m;
    *max_col = m;
    L9 r = carry_cols(t);
    for (int j = 0; j < NL; ++j) out[j] = r.l[j];
}
// product-scanning forms: mode bit 0 = p0one variant, bit 1 = digit-split instead of rippled carries
void l29_mont_ps(const Desc29* d, const uint32_t* a, const uint32_t* b, int mode, uint32_t* out, uint64_t* max_col) {
    L9 x, y;
    for (int j = 0; j < NL; ++j) { x.l[j] = a[j]; y.l[j] = b[j]; }
    uint64_t t[NL - 1];
    if (mode & 1) mont_ps_cols<true>(*d, x, y, t); else mont_ps_cols<false>(*d, x, y, t);
    uint64_t m = 0;
    for (int j = 0; j < NL - 1; ++j) m = t[j] > m ? t[j] : m;
    *max_col = m;
    L9 r = (mode & 2) ? split_cols8(t) : carry_cols8(t);
    for (int j = 0; j < NL; ++j) out[j] = r.l[j];
}
}
