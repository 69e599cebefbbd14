// Test-infrastructure only.  Drives the UNMODIFIED reference feature extraction of `filt`
// (portcullis::ml::ModelFeatures, lib/src/model_features.cc under /root/reference) and dumps the feature matrix:
//
//   feature_ref <junctions.tab> <genome.fa> <subsets.txt> <out.tsv>
//
// subsets.txt: one line per junction of the tab, three 0/1 flags "coding pass fail": which junctions train the coding-potential
// models (trainCodingPotentialModel), and which are the positive / negative sets of trainSplicingModels.  L95 comes from
// calcIntronThreshold over the `coding` subset, like JunctionFilter does with its initial positive set (src/junction_filter.cc).
// Output: one row per junction, the columns of ModelFeatures::setRow (model_features.cc:168-209) printed with 17 digits.
#include <fstream>
#include <iostream>
#include <iomanip>
#include <string>
#include <vector>
#include <portcullis/portcullis_fs.hpp>
#include <portcullis/junction_system.hpp>
#include <portcullis/ml/model_features.hpp>
#include <ranger/DataDouble.h>

portcullis::PortcullisFS portcullis::pfs;

int main(int argc, char* argv[]) {
    if (argc != 5) { std::cerr << "usage: feature_ref <junctions.tab> <genome.fa> <subsets.txt> <out.tsv>" << std::endl; return 1; }
    try {
        portcullis::JunctionSystem js; js.load(argv[1], false);
        const portcullis::JunctionList& all = js.getJunctions();
        portcullis::JunctionList coding, pass, fail;
        std::ifstream sf(argv[3]);
        for (size_t i = 0; i < all.size(); i++) { int c, p, f; if (!(sf >> c >> p >> f)) { std::cerr << "subsets.txt too short" << std::endl; return 1; }
            if (c) coding.push_back(all[i]); if (p) pass.push_back(all[i]); if (f) fail.push_back(all[i]); }
        portcullis::ml::ModelFeatures mf;
        mf.initGenomeMapper(argv[2]);
        if (!coding.empty()) { mf.calcIntronThreshold(coding); mf.trainCodingPotentialModel(coding); }
        if (!pass.empty() || !fail.empty()) mf.trainSplicingModels(pass, fail);
        Data* d = mf.juncs2FeatureVectors(all);
        std::ofstream out(argv[4]);
        out << std::setprecision(17);
        for (size_t r = 0; r < all.size(); r++) {
            for (size_t c = 0; c < d->getNumCols(); c++) out << (c ? "\t" : "") << d->get(r, c);
            out << "\n";
        }
        out << "#L95\t" << mf.L95 << "\n";
        delete d;
    }
    catch (boost::exception& e) { std::cerr << "Error: " << boost::diagnostic_information(e) << std::endl; return 4; }
    catch (std::exception& e) { std::cerr << "Error: " << e.what() << std::endl; return 5; }
    return 0;
}
