// cuhe_b200/host/cuhe_utils.cpp -- see cuhe_utils.hpp (wire format of cuhe/Utils.cu:75-152).
#include "cuhe_utils.hpp"

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

}  // namespace cuHE_Utils
