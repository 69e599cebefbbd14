// bamfilt_driver.cpp — `portcullis bamfilt`: remove the alignments that only support junctions
// missing from a junction file.  Counterpart of BamFilter::filter / main (/root/reference/src/bam_filter.cc:152-330).
//
// Host: stream the BAM (parallel inflate), write the survivors with htslib's block layout (so the output is byte-identical
// to the reference's) and index them; device: which records survive (pj_jset_filter, csrc/pj_bamfilt.cu).
#include "../../include/portcullis_junc_host.h"
#include "bam_out.hpp"
#include <algorithm>
#include <chrono>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>

namespace fs = std::filesystem;

namespace {
thread_local std::string g_berr;
int bfail(int code, const std::string& m) { g_berr = m; return code; }
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
constexpr size_t TAB_COLUMNS = 75;        // Junction::parse (junction.cc:1232-1243): 11 + 3 strands + 41 metrics + 20 JADs
}

extern "C" {

const char* pjh_bamfilt_last_error(void) { return g_berr.c_str(); }
void pjh_bamfilt_options_default(pjh_bamfilt_options* o) { memset(o, 0, sizeof *o); o->output_bam = "filtered.bam"; o->clip_mode = PJ_CLIP_HARD; o->threads = 1; }

int pjh_bamfilt_run(const pjh_bamfilt_options* o, pjh_bamfilt_report* rep) {
    if (!o || !o->junction_file || !o->bam_file || !o->output_bam) return bfail(PJ_EINVAL, "pjh_bamfilt_run: null argument");
    pjh_bamfilt_report R; memset(&R, 0, sizeof R);
    const double t0 = now_s();
    const bool say = !o->quiet;
    const int threads = std::max(1, o->threads);
    pj_jset* set = nullptr;
    try {
        if (!fs::exists(o->junction_file)) return bfail(PJ_EIO, std::string("Could not find junction file at: ") + o->junction_file);
        if (!fs::exists(o->bam_file)) return bfail(PJ_EIO, std::string("Could not find BAM file at: ") + o->bam_file);
        // ---- JunctionSystem::load (junction_system.cc:424-444): every non-empty line without "index" is a junction row ----
        if (say) std::cout << "Loading junctions from: \"" << o->junction_file << "\"" << std::endl;
        std::vector<int32_t> jt, js, je;
        {
            std::ifstream in(o->junction_file); std::string line;
            while (std::getline(in, line)) {
                size_t a = line.find_first_not_of(" \t\r\n"), z = line.find_last_not_of(" \t\r\n");
                if (a == std::string::npos) continue;
                line = line.substr(a, z - a + 1);
                if (line.find("index") != std::string::npos) continue;
                std::vector<std::string> parts; size_t p = 0;
                while (p <= line.size()) {            // boost::split with token_compress_on
                    size_t q = line.find('\t', p); if (q == std::string::npos) q = line.size();
                    if (q > p || parts.empty()) parts.push_back(line.substr(p, q - p));
                    p = q + 1;
                }
                if (parts.size() != TAB_COLUMNS)
                    return bfail(PJ_EDATA, "Could not parse line due to incorrect number of columns.  This is probably a version mismatch.  Check file and portcullis versions.  Expected "
                                           + std::to_string(TAB_COLUMNS) + " columns.  Found " + std::to_string(parts.size()) + ".");
                jt.push_back(atoi(parts[1].c_str())); js.push_back(atoi(parts[4].c_str())); je.push_back(atoi(parts[5].c_str()));
            }
        }
        R.n_junctions = (int64_t)jt.size();
        if (say) std::cout << " - Found " << jt.size() << " junctions" << std::endl << std::endl;
        {   // the device set wants (tid, start, end) order
            std::vector<size_t> ord(jt.size()); std::iota(ord.begin(), ord.end(), 0);
            std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return jt[a] != jt[b] ? jt[a] < jt[b] : js[a] != js[b] ? js[a] < js[b] : je[a] < je[b]; });
            std::vector<int32_t> t2(jt.size()), s2(jt.size()), e2(jt.size());
            for (size_t k = 0; k < ord.size(); k++) { t2[k] = jt[ord[k]]; s2[k] = js[ord[k]]; e2[k] = je[ord[k]]; }
            jt.swap(t2); js.swap(s2); je.swap(e2);
        }
        int rc = pj_jset_create(o->device, (int64_t)jt.size(), jt.data(), js.data(), je.data(), &set);
        if (rc) return bfail(rc, pj_global_last_error());

        pjio::BamFile bam; bam.open(o->bam_file);
        const std::string out_path(o->output_bam);
        {
            fs::path dir = fs::path(out_path).parent_path(); if (dir.empty()) dir = ".";
            if (!fs::exists(dir)) { if (!fs::create_directories(dir)) { pj_jset_destroy(set); return bfail(PJ_EIO, "Could not create output directory at: " + dir.string()); } }
            else if (!fs::is_directory(dir)) { pj_jset_destroy(set); return bfail(PJ_EIO, "File exists with name of suggested output directory: " + dir.string()); }
        }
        if (say) std::cout << " - Processing alignments from: \"" << o->bam_file << "\"\n - Saving filtered alignments to: \"" << out_path << "\"" << std::endl;
        const bool csi = o->use_csi != 0;
        pjio::BamOut out(out_path, true, bam.header(), csi);
        std::unique_ptr<pjio::BamOut> mod, unmod;
        // the reference always creates both MSR files' writers but only opens them with --save_msrs (bam_filter.cc:179-187)
        if (o->save_msrs) { mod = std::make_unique<pjio::BamOut>(out_path + ".mod.bam", false, bam.header(), csi); unmod = std::make_unique<pjio::BamOut>(out_path + ".unmod.bam", false, bam.header(), csi); }
        // one group of records at a time: columns for the device, raw pointers for the writers
        std::vector<const uint8_t*> ptr; std::vector<uint32_t> len, coff{0}, cig; std::vector<int32_t> tid, pos; std::vector<uint8_t> keep, nn;
        int err = PJ_OK; std::string errmsg;
        pjio::scan_records(bam, threads, [&](const uint8_t* rec, size_t l) {
            const uint8_t* r = rec + 4;
            const uint32_t l_name = r[8], n_cig = rd16(r + 12);
            ptr.push_back(rec); len.push_back((uint32_t)l); tid.push_back((int32_t)rd32(r)); pos.push_back((int32_t)rd32(r + 4));
            const uint8_t* cg = r + 32 + l_name;
            for (uint32_t k = 0; k < n_cig; k++) cig.push_back(rd32(cg + 4 * k));
            coff.push_back((uint32_t)cig.size());
        }, [&]() {
            if (err) return;
            const int64_t n = (int64_t)ptr.size();
            keep.resize((size_t)n); nn.resize((size_t)n);
            const double tg = now_s();
            err = pj_jset_filter(set, n, tid.data(), pos.data(), coff.data(), cig.data(), keep.data(), nn.data());
            R.t_device_s += now_s() - tg;
            if (err) { errmsg = pj_global_last_error(); return; }
            for (int64_t i = 0; i < n; i++) {
                R.n_in++;
                if (!keep[i]) continue;
                const uint8_t* r = ptr[i] + 4;
                const uint32_t l_name = r[8], n_cig = rd16(r + 12); const uint16_t flag = (uint16_t)rd16(r + 14);
                const uint8_t* cg = r + 32 + l_name;
                int64_t rlen = 0; for (uint32_t k = 0; k < n_cig; k++) { const uint32_t c = rd32(cg + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; }
                const bool mapped = !(flag & 0x4);
                out.add_record(ptr[i], len[i], tid[i], pos[i], (mapped && n_cig > 0) ? (int64_t)pos[i] + rlen : (int64_t)pos[i] + 1, mapped);
                R.n_out++;
                // HARD / SOFT: a kept multiply spliced read counts as modified and goes to both MSR files, unchanged (bam_filter.cc:207-219)
                if (o->clip_mode != PJ_CLIP_COMPLETE && nn[i] >= 2) {
                    R.n_modified++;
                    if (mod) { mod->add_record(ptr[i], len[i], -1, 0, 1, false); unmod->add_record(ptr[i], len[i], -1, 0, 1, false); }
                }
            }
            ptr.clear(); len.clear(); tid.clear(); pos.clear(); cig.clear(); coff.assign(1, 0);
            if (out.pending_blocks() >= 64) out.drain(threads, false);
            if (mod && mod->pending_blocks() >= 64) { mod->drain(threads, false); unmod->drain(threads, false); }
        });
        pj_jset_destroy(set); set = nullptr;
        if (err) return bfail(err, errmsg);
        out.drain(threads, true);
        if (mod) { mod->drain(threads, true); unmod->drain(threads, true); }
        if (say) std::cout << "done.\nFiltered out " << (uint32_t)(R.n_in - R.n_out) << " alignments.  In: " << R.n_in << "; Out: " << R.n_out << " (Modified: " << R.n_modified
                           << ");\n\nIndexing:\n - filtered alignments ... done." << std::endl;
    }
    catch (const std::exception& e) { if (set) pj_jset_destroy(set); return bfail(PJ_EIO, e.what()); }
    R.t_total_s = now_s() - t0;
    if (rep) *rep = R;
    return PJ_OK;
}

int pjh_bamfilt_main(int argc, char** argv) {
    pjh_bamfilt_options o; pjh_bamfilt_options_default(&o);
    std::string output = "filtered.bam", jfile, bfile, mode = "HARD"; bool help = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i]; std::string val; bool has_val = false;
        if (a.rfind("--", 0) == 0) { const size_t eq = a.find('='); if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; } }
        auto value = [&]() -> const char* { if (has_val) return val.c_str(); if (i + 1 >= argc) return nullptr; return argv[++i]; };
        if (a == "-o" || a == "--output") { const char* v = value(); if (!v) return 1; output = v; }
        else if (a == "--clip_mode") { const char* v = value(); if (!v) return 1; mode = v; }
        else if (a == "-m" || a == "--save_msrs") o.save_msrs = 1;
        else if (a == "--use_csi") o.use_csi = 1;      // the reference declares -c for both --clip_mode and --use_csi (bam_filter.cc:268-275): long forms only here
        else if (a == "-t" || a == "--threads") { const char* v = value(); if (!v) return 1; o.threads = atoi(v); }
        else if (a == "--device") { const char* v = value(); if (!v) return 1; o.device = atoi(v); }
        else if (a == "-v" || a == "--verbose") o.verbose = 1;
        else if (a == "--help") help = true;
        else if (!a.empty() && a[0] == '-' && a.size() > 1) { std::cerr << "Error: unrecognised option '" << a << "'" << std::endl; return 1; }
        else if (jfile.empty()) jfile = a; else if (bfile.empty()) bfile = a;
        else { std::cerr << "Error: too many positional options have been specified on the command line" << std::endl; return 1; }
    }
    if (help || argc <= 1) {
        std::cout << "Portcullis BAM Filter Mode Help.\n\nRemoves alignments associated with bad junctions from BAM file\n\n"
                     "Usage: portcullis bamfilt [options] <junction-file> <bam-file>\n\nOptions:\n"
                     "  -o [ --output ] arg (=\"filtered.bam\")  Output BAM file generated by this program.\n"
                     "  --clip_mode arg (=HARD)                 \"HARD\", \"SOFT\" or \"COMPLETE\" (the three modes write the same alignments, like the reference; only the Modified count differs).\n"
                     "  -m [ --save_msrs ]                      Also write the kept multiply spliced reads to <output>.mod.bam and <output>.unmod.bam.\n"
                     "  --use_csi                               Whether to use CSI indexing rather than BAI indexing.\n"
                     "  -t [ --threads ] arg (=1)               Host threads for BGZF inflate / deflate.\n"
                     "  --device arg (=0)                       The GPU that tests the alignments against the junction set.\n"
                     "  -v [ --verbose ]                        Print extra information\n  --help                                  Produce help message\n" << std::endl;
        return 1;
    }
    for (auto& ch : mode) ch = (char)toupper((unsigned char)ch);
    if (mode == "HARD") o.clip_mode = PJ_CLIP_HARD; else if (mode == "SOFT") o.clip_mode = PJ_CLIP_SOFT; else if (mode == "COMPLETE") o.clip_mode = PJ_CLIP_COMPLETE;
    else { std::cerr << "Error: Can't recognise clip mode: " << mode << std::endl; return 1; }
    o.junction_file = jfile.c_str(); o.bam_file = bfile.c_str(); o.output_bam = output.c_str();
    std::cout << "Running portcullis in BAM filter mode\n------------------------------------\n" << std::endl;
    pjh_bamfilt_report rep;
    const int rc = pjh_bamfilt_run(&o, &rep);
    if (rc) { std::cerr << "Error: " << pjh_bamfilt_last_error() << std::endl; return rc == PJ_EINVAL ? 1 : 4; }
    std::cout << "\nPortcullis BAM filter completed.\nTotal runtime: " << std::fixed << std::setprecision(1) << rep.t_total_s << "s\n" << std::endl;
    return 0;
}

} // extern "C"
