// portcullis_main.cpp — `portcullis <mode> ...` front end for the modes this repository implements.
// Mode dispatch follows src/portcullis.cc:111-127, 406-517: `prep` (alias `prepare`) and `junc` (aliases `analyse`, `analyze`),
// case-insensitive.
#include "../../include/portcullis_junc_host.h"
#include <cctype>
#include <cstring>
#include <iostream>
#include <string>

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "Usage: portcullis prep [options] <genome-file> (<bam-file>)+\n"
                     "       portcullis junc [options] <prep_data_dir>\n"
                     "       portcullis bamfilt [options] <junction-file> <bam-file>\n"
                     "This build provides `prep` (no samtools needed), the GPU `junc` stage and `bamfilt`; filt / full are the reference's own." << std::endl;
        return 1;
    }
    std::string mode(argv[1]);
    for (auto& c : mode) c = (char)std::tolower((unsigned char)c);
    if (mode == "junc" || mode == "analyse" || mode == "analyze") return pjh_junc_main(argc - 1, argv + 1);
    if (mode == "prep" || mode == "prepare") return pjh_prep_main(argc - 1, argv + 1);
    if (mode == "bamfilt" || mode == "bam_filter" || mode == "bamfilter") return pjh_bamfilt_main(argc - 1, argv + 1);
    if (mode == "--version" || mode == "-v") { std::cout << "portcullis 1.2.4 (B200 junc)" << std::endl; return 0; }
    std::cerr << "Error: mode \"" << argv[1] << "\" is not provided by this build (only `prep`, `junc` and `bamfilt`)." << std::endl;
    return 1;
}
