// mirror_host_test.cpp — the C++ mirror of the reference's interface (include/ExodusIO_b200.hpp) driven without a
// GPU (heat::Options::device = -1: file I/O and decompose only), in the reference's own call order
// (BelosMueLuSolver.cpp:165-211).  Built and run by tests/test_reference_pins.py, which then compares the file
// it wrote with what the reference's IO::decompose handed to the Exodus API.
//   mirror_host_test <mesh.exo> <out.exo> <partitions>      exit code 0 = every expectation held
#include <cstdlib>

#include "ExodusIO_b200.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 64;
    heat::Options opt;
    opt.device = -1;                                   // host-only context: no CUDA device is touched
    ExodusIO::IO io(opt);
    if (!io.open(argv[1], true)) return 1;             // ExodusIO.hpp:88
    heat::Matrix A;
    heat::Vector X, B;
    if (io.assemble(&A, &X, &B, false)) return 2;      // there is NO CPU fallback: must fail (bool false + message), not compute
    if (A || X || B) return 3;                         // outputs untouched on failure
    if (!io.create(argv[2])) return 4;                 // :103
    if (!io.decompose(std::atoi(argv[3]), false)) return 5;   // :1496
    if (io.writeSolution(heat::Vector(), 0, false)) return 6; // no vector: false, like every other failure at this boundary
    ExodusIO::IO other(opt);
    if (other.open("/nonexistent/mesh.exo", true)) return 7;  // failure = false + message on stderr (the reference: perror)
    if (other.decompose(2, false)) return 8;           // call-order contract: nothing opened (readFID == -1, :1497)
    return 0;
}
