// tests/cpp/zz_test.cpp -- CPU-only check of the NTL stand-in (cuhe_b200/host/zz_lite.hpp) that the C++ host
// layer uses when <NTL/ZZ.h> is absent: reads "a b m" triples (decimal, a and b possibly negative, m > 0) from
// stdin and prints a+b, a-b, a*b, a%m (NTL semantics: result in [0, m)), NumBits(a), and the byte round trip
// BytesFromZZ/ZZFromBytes of |a| -- compared against Python integers by tests/test_abi.py.
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "cuhe_compat.hpp"

using NTL::ZZ;

int main() {
    std::string sa, sb, sm;
    while (std::cin >> sa >> sb >> sm) {
        const ZZ a = NTL::conv<ZZ>(sa.c_str()), b = NTL::conv<ZZ>(sb.c_str()), m = NTL::conv<ZZ>(sm.c_str());
        const long nb = (NumBits(a) + 7) / 8 + 3;
        std::vector<unsigned char> buf((size_t)nb);
        BytesFromZZ(buf.data(), a, nb);
        const ZZ back = NTL::ZZFromBytes(buf.data(), nb);
        std::cout << (a + b) << ' ' << (a - b) << ' ' << (a * b) << ' ' << (a % m) << ' ' << NumBits(a) << ' ' << back
                  << ' ' << (a < b ? 1 : 0) << (a == b ? 1 : 0) << '\n';
    }
    return 0;
}
