// inflate_fast.cpp — implementation of the BGZF-block DEFLATE decoder declared in inflate_fast.hpp (RFC 1951).
#include "inflate_fast.hpp"
#include <emmintrin.h>

namespace pjinflate {

#if defined(__GNUC__)
#define PJ_ALWAYS_INLINE inline __attribute__((always_inline))
#define PJ_LIKELY(x) __builtin_expect(!!(x), 1)
#define PJ_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define PJ_ALWAYS_INLINE inline
#define PJ_LIKELY(x) (x)
#define PJ_UNLIKELY(x) (x)
#endif

uint32_t Inflater::make_entry(int kind, int sym, int consume) {
    static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (kind == 0) {
        if (sym < 256) return F_LIT | ((uint32_t)sym << 16) | (uint32_t)consume;
        if (sym == 256) return F_EOB | (uint32_t)consume;
        if (sym > 285) return 0;                                  // invalid length symbol: marks the entry unusable
        return ((uint32_t)LBASE[sym - 257] << 16) | ((uint32_t)LEXT[sym - 257] << 8) | (uint32_t)consume;
    }
    if (sym > 29) return 0;
    return ((uint32_t)DEXT[sym] << 23) | ((uint32_t)DBASE[sym] << 8) | (uint32_t)consume;   // distance: base in bits 8-22, extra count in 23-26
}

namespace {
struct Rev8 { uint8_t v[256]; Rev8() { for (int i = 0; i < 256; i++) { int r = 0; for (int k = 0; k < 8; k++) if (i & (1 << k)) r |= 0x80 >> k; v[i] = (uint8_t)r; } } };
const Rev8 REV8;
// the low l bits of c (l <= 15), reversed
inline uint32_t rev_bits(uint32_t c, int l) { return (((uint32_t)REV8.v[c & 0xff] << 8) | (uint32_t)REV8.v[(c >> 8) & 0xff]) >> (16 - l); }
}

bool Inflater::build(const uint8_t* lens, int n, int tb, uint32_t* table, int cap, int kind) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int maxlen = 15; while (maxlen > 0 && count[maxlen] == 0) maxlen--;
    if (maxlen == 0) {                                             // no codes at all: every lookup is invalid (legal for an unused distance tree)
        memset(table, 0, sizeof(uint32_t) << tb);
        return true;
    }
    // Kraft check: over-subscribed sets are rejected; incomplete sets are allowed (zlib accepts a single distance code)
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
    // canonical first codes and symbols sorted by (length, symbol)
    uint16_t next_code[17]; { uint32_t code = 0; count[0] = 0; for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = (uint16_t)code; } }
    int offs[17]; offs[1] = 0; for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
    uint16_t sorted[320];
    for (int i = 0; i < n; i++) if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t)i;
    if (left > 0) memset(table, 0, sizeof(uint32_t) << tb);        // an incomplete code leaves holes; a complete one overwrites every primary entry
    int sub_next = 1 << tb;
    int idx = 0;
    uint16_t code_of[320]; uint8_t len_of[320];
    for (int l = 1; l <= 15; l++) { uint32_t c = next_code[l]; for (int k = 0; k < count[l]; k++) { code_of[idx] = (uint16_t)c++; len_of[idx] = (uint8_t)l; idx++; } }
    // pass 1: short codes fill the primary table directly
    int i = 0;
    for (; i < idx && len_of[i] <= tb; i++) {
        const int l = len_of[i];
        const uint32_t r = rev_bits(code_of[i], l), e = make_entry(kind, sorted[i], l);
        for (uint32_t k = r; k < (1u << tb); k += (1u << l)) table[k] = e;
    }
    // pass 2: long codes, grouped by their first tb bits (consecutive in canonical order)
    while (i < idx) {
        const uint32_t prefix = (uint32_t)code_of[i] >> (len_of[i] - tb);
        int j = i, sub_bits = 0;
        while (j < idx && ((uint32_t)code_of[j] >> (len_of[j] - tb)) == prefix) { sub_bits = len_of[j] - tb; j++; }   // lengths ascend: the last is the longest
        if (sub_next + (1 << sub_bits) > cap) return false;
        const uint32_t pidx = rev_bits(prefix, tb);
        table[pidx] = F_SUB | ((uint32_t)sub_next << 16) | ((uint32_t)sub_bits << 8) | (uint32_t)tb;
        for (int k = 0; k < (1 << sub_bits); k++) table[sub_next + k] = 0;
        for (; i < j; i++) {
            const int l = len_of[i] - tb;                          // bits inside the subtable
            const uint32_t low = (uint32_t)code_of[i] & ((1u << l) - 1u);
            const uint32_t r = rev_bits(low, l), e = make_entry(kind, sorted[i], l);
            for (uint32_t k = r; k < (1u << sub_bits); k += (1u << l)) table[sub_next + k] = e;
        }
        sub_next += 1 << sub_bits;
    }
    return true;
}

// ISA is only a tag: the same source is compiled once for the baseline instruction set and once with BMI2 (shrx / bzhi make the
// variable shifts and bit-field extractions of the symbol loop single instructions without a flags dependency).
template <int ISA>
PJ_ALWAYS_INLINE bool Inflater::body(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    const uint8_t* const in_end = in + in_len;
    uint8_t* const out_begin = out; uint8_t* const out_end = out + out_len;
    uint64_t bb = 0; int bc = 0;                                   // bit buffer, bit count
    auto refill_safe = [&]() { while (bc <= 56 && in < in_end) { bb |= (uint64_t)*in++ << bc; bc += 8; } };
    auto need = [&](int n) -> bool { if (bc < n) { refill_safe(); } return bc >= n; };
    auto take = [&](int n) -> uint32_t { uint32_t v = (uint32_t)(bb & ((1ull << n) - 1ull)); bb >>= n; bc -= n; return v; };
    constexpr uint32_t LM = (1u << LT_BITS) - 1u, DM = (1u << DT_BITS) - 1u;
    for (;;) {
        if (!need(3)) return false;
        const uint32_t final_block = take(1), type = take(2);
        if (type == 0) {                                           // stored
            take(bc & 7);
            if (!need(32)) return false;
            const uint32_t len = take(16), nlen = take(16);
            if ((len ^ 0xffffu) != nlen) return false;
            // give back whole bytes still in the bit buffer
            in -= bc >> 3; bb = 0; bc = 0;
            if ((size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, in, len); in += len; out += len;
        } else if (type == 1 || type == 2) {
            const uint32_t* lt; const uint32_t* dt;
            if (type == 1) {
                if (!fixed_ready_) {
                    uint8_t l[288]; for (int k = 0; k < 144; k++) l[k] = 8; for (int k = 144; k < 256; k++) l[k] = 9; for (int k = 256; k < 280; k++) l[k] = 7; for (int k = 280; k < 288; k++) l[k] = 8;
                    uint8_t d[32]; for (int k = 0; k < 32; k++) d[k] = 5;
                    if (!build(l, 288, LT_BITS, fixed_lt_, LT_SIZE, 0) || !build(d, 32, DT_BITS, fixed_dt_, DT_SIZE, 1)) return false;
                    fixed_ready_ = true;
                }
                lt = fixed_lt_; dt = fixed_dt_;
            } else {
                if (!need(14)) return false;
                const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
                if (hlit > 286 || hdist > 30) return false;
                static const uint8_t ORD[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t cl[19] = {0};
                for (int k = 0; k < hclen; k++) { if (!need(3)) return false; cl[ORD[k]] = (uint8_t)take(3); }
                uint32_t ct[128];
                {   // plain symbol table for the code-length alphabet
                    int count[8] = {0}; for (int k = 0; k < 19; k++) count[cl[k]]++;
                    count[0] = 0; int left = 1; for (int l = 1; l <= 7; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
                    uint32_t code = 0; uint16_t nc[9]; for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; nc[l] = (uint16_t)code; }
                    for (int k = 0; k < 128; k++) ct[k] = 0;
                    for (int s = 0; s < 19; s++) if (cl[s]) {
                        const int l = cl[s]; const uint32_t r = rev_bits(nc[l]++, l);
                        for (uint32_t k = r; k < 128; k += (1u << l)) ct[k] = ((uint32_t)s << 8) | (uint32_t)l | 0x10000u;
                    }
                }
                uint8_t lens[320]; int n = 0; const int total = hlit + hdist;
                while (n < total) {
                    if (!need(7 + 7)) { refill_safe(); if (bc < 1) return false; }
                    const uint32_t e = ct[bb & 127];
                    if (!(e & 0x10000u)) return false;
                    bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                    const int s = (int)((e >> 8) & 0xff);
                    if (s < 16) lens[n++] = (uint8_t)s;
                    else {
                        int rep; uint8_t v = 0;
                        if (s == 16) { if (n == 0) return false; v = lens[n - 1]; rep = 3 + (int)take(2); }
                        else if (s == 17) rep = 3 + (int)take(3);
                        else rep = 11 + (int)take(7);
                        if (n + rep > total) return false;
                        while (rep--) lens[n++] = v;
                    }
                    if (bc < 0) return false;
                }
                if (lens[256] == 0) return false;                  // no end-of-block code
                if (!build(lens, hlit, LT_BITS, lt_, LT_SIZE, 0) || !build(lens + hlit, hdist, DT_BITS, dt_, DT_SIZE, 1)) return false;
                lt = lt_; dt = dt_;
            }
            // ---- symbol loop ----
            const uint8_t* const in_fast = (size_t)(in_end - in) > 16 ? in_end - 16 : in;          // nothing to gain (and no room) on tiny streams
            uint8_t* const out_fast = (size_t)(out_end - out) > 300 ? out_end - 300 : out;
            bool eob = false;
            for (;;) {
                if (in < in_fast && out < out_fast) {
                    // Fast loop.  One refill (>= 56 bits) covers everything one iteration consumes: three primary-table literals (<= 33 bits),
                    // or a litlen code (<= 15) + extra (<= 5) + distance code (<= 15) + extra (<= 13) = 48.  No buffer can overrun: a match
                    // is at most 258 bytes (out_fast keeps 300 of room, copies run up to 31 bytes past the match) and a refill reads 8 bytes.
                    // `e` always holds the primary entry of the next symbol, looked up BEFORE the match copy so that the load is not
                    // waiting behind it.
#define PJ_REFILL() do { uint64_t w_; memcpy(&w_, in, 8); bb |= w_ << bc; in += (63 - bc) >> 3; bc |= 56; } while (0)
                    PJ_REFILL();
                    uint32_t e = lt[bb & LM];
                    for (;;) {
                        if (e & F_LIT) {
                            bb >>= (e & 0xff); bc -= (int)(e & 0xff); *out++ = (uint8_t)(e >> 16);
                            e = lt[bb & LM];
                            if (e & F_LIT) {
                                bb >>= (e & 0xff); bc -= (int)(e & 0xff); *out++ = (uint8_t)(e >> 16);
                                e = lt[bb & LM];
                                if (e & F_LIT) { bb >>= (e & 0xff); bc -= (int)(e & 0xff); *out++ = (uint8_t)(e >> 16); e = lt[bb & LM]; }
                            }
                            if (!(in < in_fast && out < out_fast)) break;
                            PJ_REFILL();                               // the low bits e was looked up with do not change
                            continue;
                        }
                        if (PJ_UNLIKELY((e & (F_SUB | F_EOB)) != 0u || e == 0u)) {
                            if (e & F_SUB) {
                                bb >>= LT_BITS; bc -= LT_BITS;
                                e = lt[((e >> 16) & 0xfff) + (uint32_t)(bb & ((1u << ((e >> 8) & 0xf)) - 1u))];
                                if (e == 0) return false;
                                if (e & F_LIT) {
                                    bb >>= (e & 0xff); bc -= (int)(e & 0xff); *out++ = (uint8_t)(e >> 16);
                                    if (!(in < in_fast && out < out_fast)) break;
                                    PJ_REFILL(); e = lt[bb & LM];
                                    continue;
                                }
                                if (e & F_EOB) { bb >>= (e & 0xff); bc -= (int)(e & 0xff); eob = true; break; }
                                // a length code from the subtable: the match path below
                            } else if (e & F_EOB) { bb >>= (e & 0xff); bc -= (int)(e & 0xff); eob = true; break; }
                            else return false;
                        }
                        bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                        const int lx = (int)((e >> 8) & 0xff);
                        const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << lx) - 1u)); bb >>= lx; bc -= lx;
                        uint32_t d = dt[bb & DM];
                        if (PJ_UNLIKELY(d & F_SUB)) { bb >>= DT_BITS; bc -= DT_BITS; d = dt[((d >> 16) & 0xfff) + (uint32_t)(bb & ((1u << ((d >> 8) & 0xf)) - 1u))]; }
                        if (PJ_UNLIKELY(d == 0)) return false;
                        bb >>= (d & 0xff); bc -= (int)(d & 0xff);
                        const int dx = (int)((d >> 23) & 0xf);
                        const uint32_t dist = ((d >> 8) & 0x7fffu) + (uint32_t)(bb & ((1u << dx) - 1u)); bb >>= dx; bc -= dx;
                        if (PJ_UNLIKELY(dist > (size_t)(out - out_begin))) return false;
                        const bool more = in < in_fast;
                        if (more) { PJ_REFILL(); e = lt[bb & LM]; }
                        const uint8_t* src = out - dist;
                        if (PJ_LIKELY(dist >= 16)) {
                            // BAM blocks: mean match length ~30.  Two unconditional 16-byte copies, a loop only for the long matches.
                            // (the second load may overlap the first store when dist < 32: keep load / store / load / store order)
                            __m128i a = _mm_loadu_si128((const __m128i*)src); _mm_storeu_si128((__m128i*)out, a);
                            a = _mm_loadu_si128((const __m128i*)(src + 16)); _mm_storeu_si128((__m128i*)(out + 16), a);
                            if (len > 32) {
                                uint8_t* o = out + 32; const uint8_t* s = src + 32; uint8_t* const oe = out + len;
                                do { a = _mm_loadu_si128((const __m128i*)s); _mm_storeu_si128((__m128i*)o, a); s += 16; o += 16; } while (o < oe);
                            }
                        }
                        else if (dist == 1) memset(out, src[0], len);                // runs (e.g. absent base qualities, 0xff)
                        else if (dist >= 8) {
                            uint8_t* o = out; const uint8_t* s = src; uint8_t* const oe = out + len;
                            do { uint64_t w; memcpy(&w, s, 8); memcpy(o, &w, 8); s += 8; o += 8; } while (o < oe);
                        } else {
                            for (uint32_t k = 0; k < len; k++) out[k] = src[k];
                        }
                        out += len;
                        if (!(more && out < out_fast)) break;
                    }
#undef PJ_REFILL
                    if (eob) break;
                    if (bc < 0) return false;
                    continue;                                          // re-test the fast conditions; the careful step below takes over near the ends
                }
                // ---- careful single step near the ends of the buffers ----
                refill_safe();
                uint32_t e = lt[bb & LM];
                if (e & F_SUB) { bb >>= LT_BITS; bc -= LT_BITS; e = lt[(e >> 16 & 0xfff) + (bb & ((1u << ((e >> 8) & 0xf)) - 1u))]; }
                if (e == 0) return false;
                bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                if (bc < 0) return false;
                if (e & F_LIT) {
                    if (out >= out_end) return false;
                    *out++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (e & F_EOB) break;
                uint32_t len = e >> 16; const int lx = (int)((e >> 8) & 0xff);
                if (bc < lx + 28) { refill_safe(); }
                len += (uint32_t)(bb & ((1u << lx) - 1u)); bb >>= lx; bc -= lx;
                uint32_t d = dt[bb & DM];
                if (d & F_SUB) { bb >>= DT_BITS; bc -= DT_BITS; d = dt[(d >> 16 & 0xfff) + (bb & ((1u << ((d >> 8) & 0xf)) - 1u))]; }
                if (d == 0) return false;
                bb >>= (d & 0xff); bc -= (int)(d & 0xff);
                const int dx = (int)((d >> 23) & 0xf);
                const uint32_t dist = ((d >> 8) & 0x7fffu) + (uint32_t)(bb & ((1u << dx) - 1u)); bb >>= dx; bc -= dx;
                if (bc < 0) return false;
                if (dist > (size_t)(out - out_begin) || len > (size_t)(out_end - out)) return false;
                const uint8_t* src = out - dist;
                if (dist == 1) memset(out, src[0], len);
                else for (uint32_t k = 0; k < len; k++) out[k] = src[k];
                out += len;
            }
        } else return false;
        if (final_block) break;
    }
    return out == out_end;
}

bool Inflater::run_generic(Inflater* self, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) { return self->body<0>(in, in_len, out, out_len); }
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target("bmi,bmi2"))) bool Inflater::run_bmi2(Inflater* self, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) { return self->body<1>(in, in_len, out, out_len); }
static const bool HAVE_BMI2 = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("bmi");
#else
bool Inflater::run_bmi2(Inflater* self, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) { return self->body<0>(in, in_len, out, out_len); }
static const bool HAVE_BMI2 = false;
#endif

bool Inflater::run(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    return HAVE_BMI2 ? run_bmi2(this, in, in_len, out, out_len) : run_generic(this, in, in_len, out, out_len);
}

} // namespace pjinflate
