// mtx_tool -- writes a generated matrix (Poisson grids, wheel, dense, R-MAT) as a Matrix Market
// file so that the UNMODIFIED reference CG drivers, which only accept --mtx
// (cpu_singlecg.cpp:230, cpu_multicg.cpp:306), can be run on exactly the inputs the GPU sees.
// usage: mtx_tool --grid3d=150 --poisson --out=poisson150.mtx
#include "smle_host.hpp"

int main(int argc, char **argv)
{
    smle_host::Args args(argc, argv);
    std::string out;
    args.get("out", out);
    smle_host::Csr<double> a;
    std::string label = smle_host::matrix_from_args(args, a, false);
    if (label.empty() || out.empty()) { fprintf(stderr, "usage: %s <generator flags> --out=<file.mtx>\n", argv[0]); return 1; }
    if (!smle_host::write_matrix_market(out, a)) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    printf("%s: %d rows, %d nonzeros -> %s\n", label.c_str(), a.num_rows, a.num_nonzeros, out.c_str());
    return 0;
}
