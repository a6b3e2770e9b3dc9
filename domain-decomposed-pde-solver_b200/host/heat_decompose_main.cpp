// heat_decompose — CLI mirror of the reference's exec/ExodusIODecomposeTest (ExodusIODecomposeTest.cpp:5-43): same
// flags (--input --output --partitions --verbose/--no-verbose), same checks and messages, same call order
// io.open -> io.create -> io.decompose.  Pure host work (METIS + Exodus I/O): runs without a GPU.
#include <cstdlib>
#include <iostream>
#include <string>

#include "ExodusIO_b200.hpp"

int main(int argc, char *argv[]) {
    std::string inputFile, outputFile;
    size_t numPartitions = 0;
    bool verbose = false;
    heat::Options opt;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.compare(0, 8, "--input=") == 0) inputFile = a.substr(8);
        else if (a.compare(0, 9, "--output=") == 0) outputFile = a.substr(9);
        else if (a.compare(0, 13, "--partitions=") == 0) numPartitions = (size_t)std::strtoull(a.c_str() + 13, nullptr, 10);
        else if (a == "--verbose") verbose = true;
        else if (a == "--no-verbose") verbose = false;
        else if (a == "--real4") opt.output_word_size = 4;                       // float32 output like a float-real_t reference build
        else { std::cerr << "unknown option '" << a << "'" << std::endl; return EXIT_FAILURE; }
    }
    if (inputFile.empty()) {
        std::cerr << "No input file was provided; use the '--input' parameter!" << std::endl;
        return EXIT_FAILURE;
    }
    if (outputFile.empty()) {
        std::cerr << "No output file was provided; use the '--output' parameter!" << std::endl;
        return EXIT_FAILURE;
    }
    if (numPartitions == 0) {
        std::cerr << "Number of partitions to decompose the mesh has not been provided; use the '--partitions' parameter!" << std::endl;
        return EXIT_FAILURE;
    }
    opt.device = -1;                       // host-only context
    ExodusIO::IO io(opt);
    if (!io.open(inputFile, true)) {
        std::cerr << "Failed to open input Exodus file '" << inputFile << "'" << std::endl;
        return EXIT_FAILURE;
    }
    if (!io.create(outputFile)) {
        std::cerr << "Failed to create output Exodus file '" << outputFile << "'" << std::endl;
        return EXIT_FAILURE;
    }
    if (!io.decompose((int)numPartitions, verbose)) {
        std::cerr << "Failed to decompose the input file!" << std::endl;
        return EXIT_FAILURE;
    }
    return 0;
}
