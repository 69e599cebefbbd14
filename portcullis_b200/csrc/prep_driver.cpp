// prep_driver.cpp — `portcullis prep` without samtools: lays out the prep directory that `junc`
// reads.  Counterpart of Prepare::prepare / main (/root/reference/src/prepare.cc:93-148, 202-344, 384-470) and the
// PreparedFiles naming contract (src/prepare.hpp:74-140).
//
// The reference shells out to `samtools sort` / `merge` / `index` and to htslib's fai_build.  Here:
//   * genome + index        : copied or symlinked; a missing .fai is built by a restatement of fai_build_core
//                             (deps/htslib-1.3/faidx.c:82-155), byte-identical to htslib's;
//   * one sorted input BAM  : symlinked / copied like the reference;
//   * unsorted or several   : every record is read once (parallel inflate), the coordinate order comes from the GPU
//                             (pj_coordinate_order: one-sweep radix sort of samtools' key, stable, so several inputs
//                             merge exactly like `samtools merge` of their sorted versions), the records are written
//                             in that order with htslib's block layout and the header gets SO:coordinate;
//   * index                 : BAI / CSI from the same writer as `junc --separate` (csrc/bam_out.hpp).
// The sort holds the uncompressed records in host memory (the reference's samtools call works in 2 GB chunks with
// temporary files instead).
#include "../../include/portcullis_junc_host.h"
#include "bam_out.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <glob.h>
#include <climits>
#include <iomanip>
#include <iostream>
#include <unistd.h>

namespace fs = std::filesystem;

namespace {

thread_local std::string g_perr;
int pfail(int code, const std::string& m) { g_perr = m; return code; }
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
bool present(const fs::path& p) { std::error_code ec; return fs::exists(p, ec) || fs::is_symlink(fs::symlink_status(p, ec)); }

inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Prepare::copy (prepare.cc:93-124)
bool copy_or_link(const fs::path& from, const fs::path& to, const char* what, bool required, bool links, bool say) {
    if (present(to)) { if (say) std::cout << "Prepped " << what << " file detected: " << to << std::endl; }
    else if (required || present(from)) {
        if (links) { fs::create_symlink(fs::canonical(from), to); if (say) std::cout << "Created symlink from " << from << " to " << to << std::endl; }
        else {
            if (say) std::cout << "Copying from " << from << " to " << to << " ... " << std::flush;
            std::ifstream src(from.string(), std::ios::binary); std::ofstream dst(to.string(), std::ios::binary);
            dst << src.rdbuf();
            if (say) std::cout << "done." << std::endl;
        }
    }
    else if (say) std::cout << "Existing " << what << " not found.  Will create later." << std::endl;
    return present(to);
}

// fai_build_core + fai_save (htslib-1.3 faidx.c:82-175) on a plain FASTA
void build_fai(const fs::path& fasta, const fs::path& fai) {
    pjio::MappedFile f; f.open(fasta.string());
    const uint8_t* d = f.data(); const uint64_t n = f.size();
    struct Ent { std::string name; int64_t len; int line_len, line_blen; uint64_t off; };
    std::vector<Ent> ents; std::vector<std::string> seen;
    std::string name; int64_t len = -1; int line_len = -1, line_blen = -1, state = 0, l1, l2; uint64_t offset = 0, p = 0;
    auto insert = [&]() {
        if (std::find(seen.begin(), seen.end(), name) != seen.end()) { std::cerr << "[fai_build_core] ignoring duplicate sequence \"" << name << "\"" << std::endl; return; }
        seen.push_back(name); ents.push_back(Ent{name, len, line_len, line_blen, offset});
    };
    auto getc_ = [&]() -> int { return p < n ? (int)d[p++] : -1; };
    int c;
    while ((c = getc_()) >= 0) {
        if (c == '\n') {
            if (state == 1) { offset = p; continue; }
            else if ((state == 0 && len < 0) || state == 2) continue;
            else if (state == 0) { state = 2; continue; }
        }
        if (c == '>') {
            if (len >= 0) insert();
            name.clear();
            while ((c = getc_()) >= 0) { if (!isspace(c)) name.push_back((char)c); else if (!name.empty() || c == '\n') break; }
            if (c < 0) throw pjio::IoError("[fai_build_core] the last entry has no sequence");
            if (c != '\n') while ((c = getc_()) >= 0 && c != '\n') {}
            state = 1; len = 0; offset = p;
        } else {
            if (state == 3) throw pjio::IoError("[fai_build_core] inlined empty line is not allowed in sequence '" + name + "'.");
            if (state == 2) state = 3;
            l1 = l2 = 0;
            do { ++l1; if (isgraph(c)) ++l2; } while ((c = getc_()) >= 0 && c != '\n');
            if (state == 3 && l2) throw pjio::IoError("[fai_build_core] different line length in sequence '" + name + "'.");
            ++l1; len += l2;
            if (state == 1) { line_len = l1; line_blen = l2; state = 0; }
            else if (state == 0) { if (l1 != line_len || l2 != line_blen) state = 2; }
        }
    }
    if (len >= 0) insert(); else throw pjio::IoError("Genome indexing failed: " + fasta.string());
    FILE* o = fopen(fai.string().c_str(), "w");
    if (!o) throw pjio::IoError("cannot write " + fai.string());
    for (const Ent& e : ents) fprintf(o, "%s\t%d\t%lld\t%d\t%d\n", e.name.c_str(), (int)e.len, (long long)e.off, e.line_blen, e.line_len);
    fclose(o);
}

// Prepare::checkIndexMode (prepare.cc:361-382)
bool check_index_mode(const fs::path& fai, bool use_csi) {
    if (use_csi) return true;
    std::ifstream in(fai.string()); std::string line;
    while (std::getline(in, line)) {
        const size_t t1 = line.find('\t'); if (t1 == std::string::npos) continue;
        const uint64_t l = strtoull(line.c_str() + t1 + 1, nullptr, 10);
        if (l >= (uint64_t)INT32_MAX) return false;
    }
    return true;
}

// `samtools sort` rewrites the @HD line (bam_sort.c change_SO): keep the version, set SO:coordinate
std::string with_sorted_hd(const std::string& text) {
    if (text.compare(0, 3, "@HD") == 0) {
        const size_t e = text.find('\n');
        std::string hd = text.substr(0, e == std::string::npos ? text.size() : e);
        const size_t so = hd.find("\tSO:");
        if (so != std::string::npos) { const size_t q = hd.find('\t', so + 1); hd = hd.substr(0, so) + "\tSO:coordinate" + (q == std::string::npos ? "" : hd.substr(q)); }
        else hd += "\tSO:coordinate";
        return hd + (e == std::string::npos ? "\n" : text.substr(e));
    }
    return "@HD\tVN:1.3\tSO:coordinate\n" + text;
}

} // namespace

extern "C" {

const char* pjh_prep_last_error(void) { return g_perr.c_str(); }

void pjh_prep_options_default(pjh_prep_options* o) { memset(o, 0, sizeof *o); o->output_dir = "portcullis_prep"; o->threads = 1; }

int pjh_prep_run(const pjh_prep_options* o, pjh_prep_report* rep) {
    if (!o || !o->genome_file || !o->output_dir) return pfail(PJ_EINVAL, "pjh_prep_run: null argument");
    pjh_prep_report R; memset(&R, 0, sizeof R);
    const double t0 = now_s();
    const bool say = !o->quiet, links = !o->copy, csi = o->use_csi != 0;
    const int threads = std::max(1, o->threads);
    try {
        if (!present(o->genome_file)) return pfail(PJ_EIO, std::string("Could not find genome file at: ") + o->genome_file);
        if (o->n_bam_files <= 0 || !o->bam_files) return pfail(PJ_EINVAL, "No BAM files specified");
        const fs::path dir(o->output_dir);
        if (!fs::exists(dir)) { if (!fs::create_directories(dir)) return pfail(PJ_EIO, "Could not create output directory at: " + dir.string()); }
        else if (!fs::is_directory(dir)) return pfail(PJ_EIO, "File exists with name of suggested output directory: " + dir.string());
        const fs::path genome = dir / "portcullis.genome.fa", fai = dir / "portcullis.genome.fa.fai";
        const fs::path unsorted = dir / "portcullis.unsorted.alignments.bam", sorted = dir / "portcullis.sorted.alignments.bam";
        const fs::path index = fs::path(sorted.string() + (csi ? ".csi" : ".bai"));
        if (o->force) {      // PreparedFiles::clean (prepare.cc:78-87)
            if (say) std::cout << "Cleaning output dir: " << dir << " ... " << std::flush;
            for (const fs::path& p : {unsorted, sorted, fs::path(sorted.string() + ".bai"), fs::path(sorted.string() + ".csi"), genome, fai}) { std::error_code ec; fs::remove(p, ec); }
            if (say) std::cout << "done." << std::endl << std::endl;
        }
        if (!copy_or_link(o->genome_file, genome, "genome", true, links, say)) return pfail(PJ_EIO, "Could not copy/symlink genome file to: " + genome.string());
        if (!copy_or_link(std::string(o->genome_file) + ".fai", fai, "genome index", false, links, say)) {
            if (say) std::cout << "Indexing genome " << genome << " ... " << std::flush;
            build_fai(genome, fai);
            if (say) std::cout << "done." << std::endl << "Genome index file created at: " << fai << std::endl;
        }
        if (!check_index_mode(fai, csi))
            return pfail(PJ_EDATA, "User requested BAI indexing mode, however, genome file contains sequences too long to properly index using this method.  To continue, restart using the --use_csi option.");

        // ---- BAM ----
        std::vector<std::string> bams(o->bam_files, o->bam_files + o->n_bam_files);
        bool need_sort = bams.size() > 1;
        bool index_copied = false;
        if (present(sorted)) { if (say) std::cout << "Prepped sorted BAM detected: " << sorted << std::endl; need_sort = false; }
        else if (bams.size() == 1) {
            // prepare.cc:307-324: the input is first linked / copied to portcullis.unsorted.alignments.bam, then either that
            // is linked as the sorted file (already sorted) or sorted into it
            if (!copy_or_link(bams[0], unsorted, "BAM", true, links, say)) return pfail(PJ_EIO, "Could not copy/symlink BAM file to: " + unsorted.string());
            index_copied = copy_or_link(bams[0] + (csi ? ".csi" : ".bai"), index, "BAM index", false, links, say);
            pjio::BamFile probe; probe.open(unsorted.string());
            if (probe.header().text.find("SO:coordinate") != std::string::npos && !o->force) {          // BamHelper::isCoordSortedBam (bam_master.cc:46-65), prepare.cc:211
                if (links) {
                    if (say) std::cout << "Provided BAM appears to be sorted already, just creating symlink instead." << std::endl;
                    fs::create_symlink(fs::canonical(unsorted), sorted);
                } else {
                    // the reference symlinks sorted -> unsorted and then deletes the copied unsorted file (prepare.cc:321-324),
                    // which leaves a dangling link and makes its own `samtools index` fail; here the copy simply becomes the sorted file
                    if (say) std::cout << "Provided BAM appears to be sorted already: the copy becomes the sorted file." << std::endl;
                    fs::rename(unsorted, sorted);
                }
            } else { need_sort = true; bams[0] = unsorted.string(); }
        }
        if (need_sort && !present(sorted)) {
            const double ts = now_s();
            if (say) std::cout << "Sorting " << bams.size() << " BAM file(s) by coordinate (records in memory, order from the GPU) ... " << std::flush;
            // every record of every input, in input order
            std::vector<uint8_t> arena; std::vector<uint64_t> off; std::vector<int32_t> tid, pos; std::vector<uint16_t> flag;
            pjio::BamHeader hdr;
            for (size_t b = 0; b < bams.size(); b++) {
                pjio::BamFile in; in.open(bams[b]);
                if (b == 0) hdr = in.header();
                else if (in.header().names != hdr.names || in.header().lens != hdr.lens) return pfail(PJ_EDATA, "BAM files to merge have different target sequences: " + bams[b]);
                arena.reserve(arena.size() + (size_t)in.file().size() * 4);
                pjio::scan_records(in, threads, [&](const uint8_t* rec, size_t len) {
                    off.push_back(arena.size()); arena.insert(arena.end(), rec, rec + len);
                    tid.push_back((int32_t)rd32(rec + 4)); pos.push_back((int32_t)rd32(rec + 8)); flag.push_back((uint16_t)rd16(rec + 18));
                }, nullptr);
            }
            off.push_back(arena.size());
            const int64_t n = (int64_t)tid.size();
            std::vector<uint32_t> order((size_t)n);
            const double tg = now_s();
            int rc = pj_coordinate_order(o->device, n, tid.data(), pos.data(), flag.data(), order.data());
            if (rc) return pfail(rc, pj_global_last_error());
            R.t_sort_gpu_s = now_s() - tg;
            hdr.text = with_sorted_hd(hdr.text);
            // an index linked / copied from the input describes the input's block offsets, not the file written here (the
            // reference keeps it, prepare.cc:238-244, and ends up with a stale index): drop it, the writer makes a new one —
            // and must not write THROUGH a symlink into the user's own index file
            { std::error_code ec; fs::remove(index, ec); }
            pjio::BamOut out(sorted.string(), true, hdr, csi);
            for (int64_t k = 0; k < n; k++) {
                const uint32_t i = order[(size_t)k];
                const uint8_t* rec = arena.data() + off[i]; const size_t len = (size_t)(off[i + 1] - off[i]);
                const uint8_t* r = rec + 4;
                const uint32_t l_name = r[8], n_cig = rd16(r + 12); const uint8_t* cg = r + 32 + l_name;
                int64_t rlen = 0; for (uint32_t q = 0; q < n_cig; q++) { const uint32_t c = rd32(cg + 4 * q), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; }
                const bool mapped = !(flag[i] & 0x4);
                out.add_record(rec, len, tid[i], pos[i], (mapped && n_cig > 0) ? (int64_t)pos[i] + rlen : (int64_t)pos[i] + 1, mapped);
                if (out.pending_blocks() >= 256) out.drain(threads, false);
            }
            out.drain(threads, true);          // also writes the index
            R.n_records = n; R.sorted_in_process = 1; R.t_sort_s = now_s() - ts;
            // prepare.cc:321-324: save disk space by deleting a copied unsorted BAM (a symlink stays)
            { std::error_code ec; if (!fs::is_symlink(fs::symlink_status(unsorted, ec)) && fs::exists(unsorted, ec)) fs::remove(unsorted, ec); }
            if (say) std::cout << "done." << std::endl << "Sorted BAM file created at: " << sorted << std::endl << "BAM index created at: " << index << std::endl;
        }
        (void)index_copied;
        if (!present(index)) {
            // Prepare::bamIndex (prepare.cc:238-260): the reference runs `samtools index`; here one pass over the file builds it
            if (say) std::cout << "Indexing BAM ... " << std::flush;
            pjio::BamFile in; in.open(sorted.string());
            pjio::IndexBuilder ib; ib.targets.resize(in.header().lens.size());
            if (csi) { int64_t ml = 0; for (int32_t l : in.header().lens) ml = std::max<int64_t>(ml, l); ml += 256; int d = 0; for (int64_t s = 1 << 14; ml > s; ++d, s <<= 3) {} ib.depth = d; }
            pjio::index_existing_bam(in, ib);
            const std::vector<uint8_t> bytes = csi ? ib.csi() : ib.bai();
            FILE* g = fopen(index.string().c_str(), "wb");
            if (!g || fwrite(bytes.data(), 1, bytes.size(), g) != bytes.size() || fclose(g) != 0) return pfail(PJ_EIO, "Failed to successfully index: " + sorted.string());
            if (say) std::cout << "done." << std::endl << "BAM index created at: " << index << std::endl;
        }
        if (!present(sorted) || !present(index) || !present(genome) || !present(fai)) return pfail(PJ_EIO, "Prepared data is not complete: " + dir.string());
    }
    catch (const std::exception& e) { return pfail(PJ_EIO, e.what()); }
    R.t_total_s = now_s() - t0;
    if (rep) *rep = R;
    return PJ_OK;
}

static void prep_help() {
    std::cout << "Portcullis Prepare Mode Help.\n\nPrepares a genome and bam file(s) ready for junction analysis.  This involves\n"
                 "ensuring the bam file is sorted and indexed and the genome file is indexed.\n\n"
                 "Usage: portcullis prep [options] <genome-file> (<bam-file>)+\n\nOptions:\n"
                 "  -o [ --output ] arg (=portcullis_prep)  Output directory for prepared files.\n"
                 "  --force                                 Whether or not to clean the output directory before processing.\n"
                 "  --copy                                  Whether to copy files from input data to prepared data where possible, otherwise will use symlinks.\n"
                 "  -c [ --use_csi ]                        Whether to use CSI indexing rather than BAI indexing.\n"
                 "  -t [ --threads ] arg (=1)               The number of host threads used to inflate / deflate BAM blocks.\n"
                 "  --device arg (=0)                       The GPU that computes the coordinate order when a sort is needed.\n"
                 "  -v [ --verbose ]                        Print extra information\n"
                 "  --help                                  Produce help message\n" << std::endl;
}

int pjh_prep_main(int argc, char** argv) {
    pjh_prep_options o; pjh_prep_options_default(&o);
    std::string outdir = "portcullis_prep", genome; std::vector<std::string> pats; bool help = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i]; std::string val; bool has_val = false;
        if (a.rfind("--", 0) == 0) { const size_t eq = a.find('='); if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; } }
        auto value = [&]() -> const char* { if (has_val) return val.c_str(); if (i + 1 >= argc) return nullptr; return argv[++i]; };
        if (a == "-o" || a == "--output") { const char* v = value(); if (!v) { std::cerr << "Error: the required argument for option '--output' is missing" << std::endl; return 1; } outdir = v; }
        else if (a == "-t" || a == "--threads") { const char* v = value(); if (!v) return 1; o.threads = atoi(v); }
        else if (a == "--device") { const char* v = value(); if (!v) return 1; o.device = atoi(v); }
        else if (a == "--force") o.force = 1;
        else if (a == "--copy") o.copy = 1;
        else if (a == "-c" || a == "--use_csi") o.use_csi = 1;
        else if (a == "-v" || a == "--verbose") o.verbose = 1;
        else if (a == "--help") help = true;
        else if (!a.empty() && a[0] == '-' && a.size() > 1) { std::cerr << "Error: unrecognised option '" << a << "'" << std::endl; return 1; }
        else if (genome.empty()) genome = a; else pats.push_back(a);
    }
    if (help || argc <= 1) { prep_help(); return 1; }
    // Prepare::globFiles (prepare.cc:346-359)
    std::vector<std::string> bams;
    for (const std::string& p : pats) { glob_t g; if (glob(p.c_str(), 0, nullptr, &g) == 0) { for (size_t k = 0; k < g.gl_pathc; k++) bams.push_back(g.gl_pathv[k]); } globfree(&g); }
    std::vector<const char*> ptrs; for (auto& b : bams) ptrs.push_back(b.c_str());
    o.genome_file = genome.c_str(); o.bam_files = ptrs.data(); o.n_bam_files = (int32_t)ptrs.size(); o.output_dir = outdir.c_str();
    std::cout << "Running portcullis in prepare mode\n----------------------------------\n" << std::endl;
    pjh_prep_report rep;
    const int rc = pjh_prep_run(&o, &rep);
    if (rc) { std::cerr << "Error: " << pjh_prep_last_error() << std::endl; return rc == PJ_EINVAL ? 1 : 4; }
    std::cout << "\nPortcullis prep completed.\nTotal runtime: " << std::fixed << std::setprecision(1) << rep.t_total_s << "s";
    if (rep.sorted_in_process) std::cout << "  (" << rep.n_records << " records sorted; coordinate order on the GPU " << std::setprecision(3) << rep.t_sort_gpu_s << "s)";
    std::cout << "\n" << std::endl;
    return 0;
}

} // extern "C"
