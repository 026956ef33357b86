// cuhe_b200/host/cuhe_utils.cpp -- see cuhe_utils.hpp (wire format of cuhe/Utils.cu:75-152).
#include "cuhe_utils.hpp"

#include <cstdio>
#include <cstring>
#include <stdexcept>

#include <sstream>

namespace cuHE_Utils {

namespace {
// tokens between any of the characters of `delims`; empty tokens are dropped (the reference
// tokenises with strtok_r, cuhe/Utils.cu:77-93, which behaves this way)
std::vector<std::string> tokens(const std::string& s, const std::string& delims) {
    std::vector<std::string> out;
    size_t pos = 0;
    while (pos < s.size()) {
        const size_t start = s.find_first_not_of(delims, pos);
        if (start == std::string::npos) break;
        size_t end = s.find_first_of(delims, start);
        if (end == std::string::npos) end = s.size();
        out.emplace_back(s, start, end - start);
        pos = end;
    }
    return out;
}
std::string decimal(const ZZ& v) {
    std::ostringstream os;
    os << v;
    return os.str();
}
}  // namespace

Picklable::Picklable(std::string key, ZZ* coeffs, int len) : key_(key), sep_(",") {
    coeffs_.assign(coeffs, coeffs + (len > 0 ? len : 0));
    for (int i = 0; i < len; i++) SetCoeff(poly_, i, coeffs[i]);
    render();
}
Picklable::Picklable(std::string key, ZZX poly) : key_(key), sep_(","), poly_(poly) {
    for (long i = 0; i <= deg(poly_); i++) coeffs_.push_back(coeff(poly_, i));
    render();
}
Picklable::Picklable(std::string data) : sep_(",") { parse(data); }
Picklable::Picklable(std::string data, std::string sep) : sep_(sep) { parse(data); }
Picklable::Picklable(const Picklable& o) : key_(o.key_), sep_(o.sep_), poly_(o.poly_) {
    for (long i = 0; i <= deg(poly_); i++) coeffs_.push_back(coeff(poly_, i));   // cuhe/Utils.cu:64-72
    render();
}
Picklable& Picklable::operator=(const Picklable& o) {
    if (this != &o) { key_ = o.key_; sep_ = o.sep_; poly_ = o.poly_; coeffs_ = o.coeffs_; values_ = o.values_; }
    return *this;
}
Picklable::~Picklable() {}

void Picklable::parse(const std::string& data) {
    const std::vector<std::string> t = tokens(data, sep_);
    clear(poly_);
    coeffs_.clear();
    if (!t.empty()) key_ = t[0];
    for (size_t i = 1; i < t.size(); i++) SetCoeff(poly_, (long)i - 1, NTL::conv<ZZ>(t[i].c_str()));
    for (long i = 0; i <= deg(poly_); i++) coeffs_.push_back(coeff(poly_, i));
    render();
}
void Picklable::render() {
    std::string s;
    for (size_t i = 0; i < coeffs_.size(); i++) {
        if (i) s += sep_;
        s += decimal(coeffs_[i]);
    }
    values_ = s;
}
void Picklable::setSeparator(std::string sep) { sep_ = sep; render(); }
std::string Picklable::getSeparator() { return sep_; }
ZZX Picklable::getPoly() { return poly_; }
ZZ* Picklable::getCoeffs() { return coeffs_.empty() ? nullptr : coeffs_.data(); }
int Picklable::getCoeffsLen() { return (int)coeffs_.size(); }
std::string Picklable::getKey() { return key_; }
std::string Picklable::getValues() { return values_; }
std::string Picklable::pickle() { return key_ + sep_ + values_; }

PicklableMap::PicklableMap(std::vector<Picklable*> items) : items_(items), sep_("\n") {}
PicklableMap::PicklableMap(std::string data) : sep_("\n") { parse(data, ","); }
PicklableMap::PicklableMap(std::string data, std::string field_sep) : sep_("\n") { parse(data, field_sep); }
PicklableMap::PicklableMap(std::string data, std::string record_sep, std::string field_sep) : sep_(record_sep) {
    parse(data, field_sep);
}
PicklableMap::~PicklableMap() {
    for (Picklable* p : owned_) delete p;
}
void PicklableMap::parse(const std::string& data, const std::string& field_sep) {
    for (Picklable* p : owned_) delete p;
    owned_.clear();
    items_.clear();
    for (const std::string& rec : tokens(data, sep_)) {
        owned_.push_back(new Picklable(rec, field_sep));
        items_.push_back(owned_.back());
    }
}
void PicklableMap::setSeparator(std::string sep) { sep_ = sep; }
std::string PicklableMap::getSeparator() { return sep_; }
std::vector<Picklable*> PicklableMap::getPicklables() { return items_; }
std::string PicklableMap::toString() {
    std::string s;
    for (size_t i = 0; i < items_.size(); i++) {
        if (i) s += sep_;
        s += items_[i]->pickle();
    }
    return s;
}
Picklable* PicklableMap::get(std::string key) {
    for (Picklable* p : items_)
        if (p->getKey() == key) return p;
    throw "not found";
}

// ---- binary RNS container ---------------------------------------------------------------------------------------
uint64_t RnsBlob::checksum(const unsigned char* data, size_t bytes) {
    uint64_t h = 0xCBF29CE484222325ull;
    const uint64_t prime = 0x100000001B3ull;
    const size_t words = (bytes + 7) / 8, block = (size_t)1 << 16;
    for (size_t off = 0; off < words; off += block) {
        uint64_t x = 0;
        const size_t end = off + block < words ? off + block : words;
        for (size_t i = off; i < end; i++) {
            uint64_t w = 0;
            const size_t take = bytes - i * 8 < 8 ? bytes - i * 8 : 8;
            std::memcpy(&w, data + i * 8, take);          // little-endian host
            x ^= w * (((uint64_t)i << 1) | 1);
        }
        h = (h ^ x) * prime;
    }
    return h;
}
namespace {
struct RnsHeader {                                        // 80 bytes, little-endian
    char magic[8];
    int32_t params[6];
    int32_t domain, level, shard_rank, shard_world;
    uint32_t ndim, dims[3];
    uint64_t bytes, digest;
};
static_assert(sizeof(RnsHeader) == 80, "RNS header layout");
}  // namespace
void RnsBlob::save(const std::string& path) const {
    if (dims.empty() || dims.size() > 3) throw std::runtime_error("RnsBlob: 1 to 3 dimensions");
    RnsHeader h{};
    std::memcpy(h.magic, "CUHERNS1", 8);
    for (int i = 0; i < 6; i++) h.params[i] = params[i];
    h.domain = domain; h.level = level; h.shard_rank = shard_rank; h.shard_world = shard_world;
    h.ndim = (uint32_t)dims.size();
    for (int i = 0; i < 3; i++) h.dims[i] = i < (int)dims.size() ? dims[(size_t)i] : 1u;
    h.bytes = payload.size();
    h.digest = checksum(payload.data(), payload.size());
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("RnsBlob: cannot open " + path);
    const bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && (payload.empty() || std::fwrite(payload.data(), payload.size(), 1, f) == 1);
    std::fclose(f);
    if (!ok) throw std::runtime_error("RnsBlob: short write to " + path);
}
RnsBlob RnsBlob::load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("RnsBlob: cannot open " + path);
    RnsHeader h{};
    RnsBlob b;
    bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "CUHERNS1", 8) == 0 && h.ndim >= 1 && h.ndim <= 3;
    if (ok) {
        b.payload.resize((size_t)h.bytes);
        ok = h.bytes == 0 || std::fread(b.payload.data(), (size_t)h.bytes, 1, f) == 1;
        unsigned char extra;
        ok = ok && std::fread(&extra, 1, 1, f) == 0;
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("RnsBlob: not a CUHERNS1 file or truncated: " + path);
    if (checksum(b.payload.data(), b.payload.size()) != h.digest) throw std::runtime_error("RnsBlob: checksum mismatch in " + path);
    for (int i = 0; i < 6; i++) b.params[i] = h.params[i];
    b.domain = h.domain; b.level = h.level; b.shard_rank = h.shard_rank; b.shard_world = h.shard_world;
    b.dims.assign(h.dims, h.dims + h.ndim);
    const size_t elem = b.domain == 3 ? 8 : 4;
    size_t count = 1;
    for (uint32_t d : b.dims) count *= d;
    if (count * elem != b.payload.size()) throw std::runtime_error("RnsBlob: extents do not match the payload in " + path);
    return b;
}

}  // namespace cuHE_Utils
