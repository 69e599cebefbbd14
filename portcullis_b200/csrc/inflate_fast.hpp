// inflate_fast.hpp — a small, fast raw-DEFLATE (RFC 1951) decoder for BGZF blocks (implementation: inflate_fast.cpp).
//
// BGZF decode is the host-side bottleneck of `junc` (inflate is ~75% of the decode CPU time), and every BGZF block
// is at most 64 KiB with its exact output size stored in the trailer.  This decoder is written for that case: 64-bit bit
// buffer refilled with unaligned 8-byte loads, 11-bit primary litlen table (8-bit for distances) with subtables for longer
// codes, a fast loop that needs no bounds checks away from the buffer ends and looks the next symbol up before it copies the
// current match, 16-byte match copies, and a BMI2 build of the same loop picked at run time.  Any stream it does not like
// (malformed code lengths, table overflow, size mismatch) makes it return false and the caller falls back to zlib, so it can
// only ever be an accelerator.
#pragma once
#include <cstdint>
#include <cstring>

namespace pjinflate {

class Inflater {
public:
    // Decodes one complete raw deflate stream of `in_len` bytes into exactly `out_len` bytes.
    // `out` must have at least out_len + 16 bytes of capacity.  Returns false if the stream is not accepted.
    bool run(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);

private:
    static constexpr int LT_BITS = 11, DT_BITS = 8;
    static constexpr uint32_t F_LIT = 0x80000000u, F_EOB = 0x40000000u, F_SUB = 0x20000000u;
    static constexpr int LT_SIZE = 4096, DT_SIZE = 1024;
    uint32_t lt_[LT_SIZE];
    uint32_t dt_[DT_SIZE];
    bool fixed_ready_ = false;
    uint32_t fixed_lt_[LT_SIZE];
    uint32_t fixed_dt_[DT_SIZE];

    // Builds a decode table from code lengths.  kind 0 = litlen, 1 = distance.
    static bool build(const uint8_t* lens, int n, int table_bits, uint32_t* table, int table_cap, int kind);
    static uint32_t make_entry(int kind, int sym, int consume);
    template <int ISA> bool body(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
    static bool run_generic(Inflater* self, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
    static bool run_bmi2(Inflater* self, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
};

} // namespace pjinflate
