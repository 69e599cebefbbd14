// Test-infrastructure only. Tiny driver that exposes the UNMODIFIED reference `prep`, `junc` and `bamfilt`
// stages (src/prepare.cc, src/junction_builder.cc, src/bam_filter.cc under /root/reference) as a CLI, without the
// reference's src/portcullis.cc (which drags in filter / ranger / embedded CPython).
// Mirrors the dispatch in /root/reference/src/portcullis.cc:406-517 for these two modes only.
#include <iostream>
#include <string>
#include <cstring>
#include <portcullis/portcullis_fs.hpp>
#include <portcullis/junction_system.hpp>
#include "junction_builder.hpp"
#include "prepare.hpp"
#include "bam_filter.hpp"

portcullis::PortcullisFS portcullis::pfs;

int main(int argc, char* argv[]) {
    if (argc < 2) { std::cerr << "usage: portcullis_ref prep|junc|bamfilt [options]" << std::endl; return 1; }
    portcullis::JunctionSystem::version = "1.2.4";   // PACKAGE_VERSION, configure.ac:7
    std::string mode(argv[1]);
    try {
        if (mode == "prep") return portcullis::Prepare::main(argc - 1, argv + 1);
        if (mode == "junc") return portcullis::JunctionBuilder::main(argc - 1, argv + 1);
        if (mode == "bamfilt") return portcullis::BamFilter::main(argc - 1, argv + 1);
        std::cerr << "unknown mode " << mode << std::endl; return 1;
    }
    catch (boost::exception& e) { std::cerr << "Error: " << boost::diagnostic_information(e) << std::endl; return 4; }
    catch (std::exception& e) { std::cerr << "Error: " << e.what() << std::endl; return 5; }
    catch (...) { std::cerr << "Error: unknown" << std::endl; return 7; }
}
