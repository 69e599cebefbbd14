// ASAN/UBSAN fuzz of the BGZF inflater: real BGZF blocks of a BAM, bit flips / truncated input / wrong expected size; what it accepts must equal zlib.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -Iportcullis_b200/csrc -o /tmp/inflate_fuzz tools/inflate_fuzz.cpp portcullis_b200/csrc/inflate_fast.cpp -lz
//   ASAN_OPTIONS=detect_leaks=0 /tmp/inflate_fuzz <file.bam> <MB to read> <iterations>
#include "inflate_fast.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <zlib.h>
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); size_t n = (size_t)atol(argv[2]) << 20;
    std::vector<uint8_t> buf(n); n = fread(buf.data(), 1, n, f); fclose(f);
    struct Blk { size_t off, clen, isize; }; std::vector<Blk> blks;
    for (size_t p = 0; p + 18 <= n;) { const uint8_t* h = &buf[p]; uint32_t xlen = h[10] | (h[11] << 8); uint32_t bsize = (h[16] | (h[17] << 8)) + 1u; if (p + bsize > n) break;
        uint32_t isize; memcpy(&isize, h + bsize - 4, 4); blks.push_back({p + 12 + xlen, bsize - 12 - xlen - 8, isize}); p += bsize; }
    pjinflate::Inflater inf; uint64_t x = 88172645463325252ull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long ok = 0, rej = 0, wrong = 0;
    z_stream z; memset(&z, 0, sizeof z); inflateInit2(&z, -15);
    for (int it = 0; it < atoi(argv[3]); it++) {
        const Blk& b = blks[rnd() % blks.size()];
        // exact-size heap copies so that ASAN sees any over-read / over-write beyond the documented slack (out: +16)
        std::vector<uint8_t> in(buf.begin() + b.off, buf.begin() + b.off + b.clen);
        const int mode = (int)(rnd() % 4);
        size_t in_len = in.size(), out_len = b.isize;
        if (mode == 1) for (int k = 0; k < 3; k++) in[rnd() % in.size()] ^= (uint8_t)(1u << (rnd() % 8));      // bit flips
        if (mode == 2) in_len = rnd() % in.size();                                                              // truncated input
        if (mode == 3) out_len = b.isize > 10 ? b.isize - 1 - rnd() % 10 : b.isize;                             // wrong expected size
        std::vector<uint8_t> in2(in.begin(), in.begin() + in_len);
        std::vector<uint8_t> out(out_len + 16);
        const bool r = inf.run(in2.data(), in2.size(), out.data(), out_len);
        if (!r) { rej++; continue; }
        // accepted: zlib must agree byte for byte
        std::vector<uint8_t> ref(out_len + 16);
        inflateReset(&z); z.next_in = in2.data(); z.avail_in = (uInt)in2.size(); z.next_out = ref.data(); z.avail_out = (uInt)out_len;
        const int zr = inflate(&z, Z_FINISH);
        if ((zr == Z_STREAM_END || zr == Z_OK || zr == Z_BUF_ERROR) && z.total_out == out_len && memcmp(ref.data(), out.data(), out_len) == 0) ok++; else wrong++;
    }
    printf("accepted and equal to zlib %ld, rejected %ld, accepted but different %ld\n", ok, rej, wrong);
    return wrong != 0;
}
