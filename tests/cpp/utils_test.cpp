// tests/cpp/utils_test.cpp -- CPU-only client of cuhe_b200/host/cuhe_utils.hpp: the key/polynomial
// text format of the reference (cuhe/Utils.h:39-93; used by examples/DHS/DHS.cu:57-189).
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cuhe_utils.hpp"

using namespace cuHE_Utils;
using NTL::ZZ;
using NTL::ZZX;

static int fails = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL line %d: %s\n", __LINE__, #cond); fails++; } } while (0)

int main(int argc, char** argv) {
    // decimal text of big integers round-trips (needed by the format)
    const char* big = "1234567890123456789012345678901234567890123456789012345678901234567890";
    ZZ b = NTL::conv<ZZ>(big);
    {
        std::ostringstream os; os << b;
        EXPECT(os.str() == big);
        std::ostringstream os2; os2 << (ZZ(0) - b);
        EXPECT(os2.str() == std::string("-") + big);
        std::ostringstream os3; os3 << ZZ(0) << "," << ZZ(1000000000L) << "," << ZZ(-7);
        EXPECT(os3.str() == "0,1000000000,-7");
        EXPECT(b % ZZ(1000000007L) == NTL::conv<ZZ>("1234567890123456789012345678901234567890123456789012345678901234567890") % ZZ(1000000007L));
    }
    // from a ZZX: deg+1 coefficients, ascending, comma separated, key first
    ZZX p;
    SetCoeff(p, 0, 5); SetCoeff(p, 1, b); SetCoeff(p, 3, 7);
    Picklable pk("pk0", p);
    EXPECT(pk.getKey() == "pk0" && pk.getCoeffsLen() == 4);
    EXPECT(pk.getValues() == std::string("5,") + big + ",0,7");
    EXPECT(pk.pickle() == std::string("pk0,5,") + big + ",0,7");
    // from a coefficient array: the given length is kept, trailing zeros included (coeffMod, d, p, ...)
    ZZ arr[3] = {ZZ(24), ZZ(0), ZZ(0)};
    Picklable d("d", arr, 3);
    EXPECT(d.pickle() == "d,24,0,0" && d.getCoeffsLen() == 3 && deg(d.getPoly()) == 0);
    // parsing, separators, copy
    Picklable q(pk.pickle());
    EXPECT(q.getKey() == "pk0" && q.getPoly() == p && q.getCoeffs()[1] == b);
    Picklable semi("x;1;;2", ";");                       // empty fields are skipped, as strtok does
    EXPECT(semi.getCoeffsLen() == 2 && semi.pickle() == "x;1;2");
    semi.setSeparator(":");
    EXPECT(semi.pickle() == "x:1:2");
    Picklable cp(pk);
    EXPECT(cp.pickle() == pk.pickle());
    // map: records joined by "\n", lookup by key, "not found" is thrown as a C string
    std::vector<Picklable*> items = {&d, &pk};
    PicklableMap m(items);
    const std::string text = m.toString();
    EXPECT(text == d.pickle() + "\n" + pk.pickle());
    PicklableMap parsed(text);
    EXPECT(parsed.getPicklables().size() == 2);
    EXPECT(parsed.get("pk0")->getPoly() == p && parsed.get("d")->getValues() == "24");   // parsing drops trailing zeros
    bool threw = false;
    try { parsed.get("nope"); } catch (const char* e) { threw = std::string(e) == "not found"; }
    EXPECT(threw);
    PicklableMap custom("a:1:2|b:3", "|", ":");
    EXPECT(custom.getPicklables().size() == 2 && custom.get("b")->getValues() == "3" && custom.toString() == "a:1:2|b:3");
    // binary RNS container: write, read back, detect damage; with a path argument: read a file written by
    // cuhe_b200/utils.py save_rns, check it, and write it back out for a byte comparison on the Python side
    {
        RnsBlob blob;
        const int ps[6] = {3, 2, 16, 48, 24, 32767};
        for (int i = 0; i < 6; i++) blob.params[i] = ps[i];
        blob.domain = 3; blob.level = 1; blob.shard_rank = 2; blob.shard_world = 4;
        blob.dims = {2, 3, 5};
        blob.payload.resize(2 * 3 * 5 * 8);
        for (size_t i = 0; i < blob.payload.size(); i++) blob.payload[i] = (unsigned char)(i * 37 + 11);
        const std::string path = argc > 1 ? std::string(argv[1]) + ".cpp_out" : std::string("/tmp/cuhe_b200_utils_test.rns");
        blob.save(path);
        RnsBlob back = RnsBlob::load(path);
        EXPECT(back.payload == blob.payload && back.dims == blob.dims && back.domain == 3 && back.level == 1 &&
               back.shard_rank == 2 && back.shard_world == 4 && back.params[5] == 32767);
        FILE* f = std::fopen(path.c_str(), "r+b");
        std::fseek(f, 100, SEEK_SET);
        std::fputc(0x5a, f);
        std::fclose(f);
        bool damaged = false;
        try { RnsBlob::load(path); } catch (const std::runtime_error&) { damaged = true; }
        EXPECT(damaged);
        std::remove(path.c_str());
        if (argc > 1) {
            RnsBlob py = RnsBlob::load(argv[1]);
            EXPECT(py.domain == 2 && py.dims.size() == 2);
            py.save(std::string(argv[1]) + ".roundtrip");
        }
    }
    if (fails == 0) std::printf("utils ok\n");
    return fails ? 1 : 0;
}
