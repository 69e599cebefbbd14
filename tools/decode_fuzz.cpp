// ASAN/UBSAN fuzz of the BAM record parser behind the lean decode (bam_io.cpp): the uncompressed stream of a small BAM is mutated
// (byte flips, length fields, truncation), re-framed as valid BGZF and decoded in both batch forms; errors must come out as exceptions.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread -Iportcullis_b200/csrc -Iinclude -o /tmp/decode_fuzz tools/decode_fuzz.cpp \
//       portcullis_b200/csrc/bam_io.cpp portcullis_b200/csrc/inflate_fast.cpp portcullis_b200/csrc/fasta_io.cpp -lz
//   ASAN_OPTIONS=detect_leaks=0 /tmp/decode_fuzz tests/golden/kat/reads.bam 2000
#include "bam_io.hpp"
#include "bam_write.hpp"
#include "fasta_io.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>
using namespace pjio;
int main(int argc, char** argv) {
    MappedFile mf; mf.open(argv[1]);
    std::vector<uint8_t> raw;
    { BgzfStream s(mf); s.seek(0); std::vector<uint8_t> b(1 << 16); size_t g; while ((g = s.read(b.data(), b.size()))) raw.insert(raw.end(), b.begin(), b.begin() + g); }
    BamFile ref; ref.open(argv[1]);
    const size_t body = (size_t)(ref.header().first_record_voff >> 16) ? 0 : (size_t)(ref.header().first_record_voff & 0xffff);   // records start here when the header fits one block
    uint64_t x = 0x2545F4914F6CDD1Dull; auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long ok = 0, thrown = 0; const int iters = atoi(argv[2]);
    const std::string tmp = "/tmp/decode_fuzz_" + std::to_string(getpid()) + ".bam";
    for (int it = 0; it < iters; it++) {
        std::vector<uint8_t> m = raw;
        const int mode = (int)(rnd() % 5);
        const size_t lo = body ? body : m.size() / 4;
        if (mode <= 1) for (int k = 0; k < 1 + (int)(rnd() % 6); k++) m[lo + rnd() % (m.size() - lo)] ^= (uint8_t)(1u << (rnd() % 8));
        if (mode == 2) for (int k = 0; k < 3; k++) m[lo + rnd() % (m.size() - lo)] = (uint8_t)rnd();
        if (mode == 3) m.resize(lo + rnd() % (m.size() - lo));
        if (mode == 4) { size_t p = lo + rnd() % (m.size() - lo - 8); uint32_t v = (uint32_t)rnd() >> (rnd() % 32); memcpy(&m[p], &v, 4); }
        std::vector<uint8_t> out;
        for (size_t o = 0; o < m.size(); o += 60000) bamw::bgzf_block(m.data() + o, std::min<size_t>(60000, m.size() - o), out, 1);
        bamw::bgzf_eof(out);
        { std::ofstream f(tmp, std::ios::binary); f.write((const char*)out.data(), (std::streamsize)out.size()); }
        try {
            BamFile f; f.open(tmp);
            for (int lean = 0; lean < 2; lean++) for (int names = 0; names < 2; names++) {
                ColumnarChunk c; c.lean = lean != 0; c.keep_mate = (it & 1) != 0; c.with_names = names != 0;
                DecodeTask t = f.whole_file_task();
                if (lean) { t.tid = (int32_t)(rnd() % std::max<size_t>(f.header().lens.size(), 1)); }
                f.decode(t, c);
                if (c.lean && ((size_t)c.n() != c.flag.size() || c.n_cigar.size() != (size_t)c.n())) { printf("inconsistent column sizes\n"); return 1; }
            }
            ok++;
        } catch (const std::exception&) { thrown++; }
    }
    remove(tmp.c_str());
    printf("%ld decoded, %ld rejected with an exception\n", ok, thrown);
    return 0;
}
