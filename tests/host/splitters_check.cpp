// CPU check of cbl_b200/csrc/splitters.hpp (the splitters of cbl_create_sharded): reads a sample of u32 prefixes,
// prints the world - 1 splitters.   splitters_check <sample.bin> <world>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../cbl_b200/csrc/splitters.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint32_t> pre;
    uint32_t v;
    while (std::fread(&v, 4, 1, f) == 1) pre.push_back(v);
    std::fclose(f);
    for (uint32_t s : cbl::equal_cost_splitters(pre, std::atoi(argv[2]))) std::printf("%u\n", s);
    return 0;
}
