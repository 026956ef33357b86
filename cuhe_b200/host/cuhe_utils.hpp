// cuhe_b200/host/cuhe_utils.hpp
// Key / polynomial wire format of the reference (namespace cuHE_Utils, cuhe/Utils.h:39-93): one
// record per polynomial, `key<sep>c0<sep>c1...` with decimal coefficients in ascending order
// (default separator ","), records joined by "\n" in a PicklableMap.  Same class names, constructors
// and accessors, so examples/DHS/DHS.cu:57-189 (key import/export) compiles against this header.
// Pure host code; ZZ/ZZX are NTL's when present, zz_lite's otherwise.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cuhe_compat.hpp"

namespace cuHE_Utils {

using NTL::ZZ;
using NTL::ZZX;

class Picklable {
public:
    Picklable(std::string key, ZZ* coeffs, int len);        // len coefficients, trailing zeros kept
    Picklable(std::string key, ZZX poly);                   // deg(poly)+1 coefficients
    Picklable(std::string data);                            // parse "key,c0,c1,..."
    Picklable(std::string data, std::string sep);
    Picklable(const Picklable& other);
    Picklable& operator=(const Picklable& other);
    ~Picklable();

    void setSeparator(std::string sep);
    std::string getSeparator();
    ZZX getPoly();
    ZZ* getCoeffs();                                        // owned by the object
    int getCoeffsLen();
    std::string getKey();
    std::string getValues();                                // "c0,c1,..." without the key
    std::string pickle();                                   // "key,c0,c1,..."

private:
    void parse(const std::string& data);
    void render();
    std::string key_, values_, sep_;
    ZZX poly_;
    std::vector<ZZ> coeffs_;
};

class PicklableMap {
public:
    PicklableMap(std::vector<Picklable*> items);            // takes the pointers as they are (not owned)
    PicklableMap(std::string data);                         // records split on "\n", fields on ","
    PicklableMap(std::string data, std::string field_sep);
    PicklableMap(std::string data, std::string record_sep, std::string field_sep);
    ~PicklableMap();

    void setSeparator(std::string sep);
    std::string getSeparator();
    std::vector<Picklable*> getPicklables();
    std::string toString();
    Picklable* get(std::string key);                        // throws (const char*) "not found"

private:
    void parse(const std::string& data, const std::string& field_sep);
    std::vector<Picklable*> items_;
    std::vector<Picklable*> owned_;                         // the ones parse() created
    std::string sep_;
};

// ---- binary RNS container (no reference counterpart; cuhe/Utils.cu:75-152 only has the decimal text form) ------------
// Residue-domain data exactly as the device holds it -- CRT domain u32[rows][crtLen] or NTT domain
// u64[rows][..][nttLen], e.g. the transformed evaluation keys of cuhe_relin_export_host / cuhe_relin_import_host --
// with the setParameters tuple that fixes its meaning and a checksum.  Same bytes as cuhe_b200/utils.py save_rns /
// load_rns (80-byte little-endian header "CUHERNS1", see there); files are interchangeable between the two.
struct RnsBlob {
    int params[6] = {0, 0, 0, 0, 0, 0};        // d, p, w, min, cut, m
    int domain = 2;                              // 2 = CRT (u32 elements), 3 = NTT (u64 elements)
    int level = 0, shard_rank = 0, shard_world = 1;
    std::vector<uint32_t> dims;                  // 1 to 3 extents
    std::vector<unsigned char> payload;          // little-endian elements

    static uint64_t checksum(const unsigned char* data, size_t bytes);
    void save(const std::string& path) const;    // throws std::runtime_error
    static RnsBlob load(const std::string& path);  // throws std::runtime_error on a damaged file
};

}  // namespace cuHE_Utils
