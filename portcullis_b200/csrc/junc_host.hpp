// junc_host.hpp — host-side tail of the junc path: merge/sort/index/group statistics (A12/A13) and the
// byte-exact junctions.tab / .bed / .gff3 writers (A14).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/portcullis_junc.h"

namespace pjhost {

struct TargetInfo { std::string name; int32_t length; };

// JunctionSystem::saveAll (lib/src/junction_system.cc:336-383): <prefix>.junctions.tab, .bed and optional GFFs.
// `version` is the string of the BED track line (JunctionSystem::version; "X.X.X" when empty).
// extra: the `--extra` columns (mm_score, coverage, up_aln, down_aln) per row, or nullptr to print them as 0 like plain `junc`
void write_tab(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets, const pj_junction_extra* extra = nullptr);
void write_bed(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
               const std::string& source, const std::string& version);
void write_exon_gff(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
                    const std::string& source);
void write_intron_gff(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
                      const std::string& source);
std::string tab_header();
int format_selftest(int n_cases);      // writers' number formatting against printf; number of differences

} // namespace pjhost
