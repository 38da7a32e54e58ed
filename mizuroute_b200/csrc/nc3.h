// Minimal NetCDF-3 (classic CDF-1 and 64-bit-offset CDF-2) reader and writer.
//
// The reference reads its river-network and runoff files through netCDF-Fortran (read_streamSeg.f90:44,
// read_runoff.f90, ncio_utils.f90) and writes history files through PIO.  Neither library exists in this image,
// so the stand-alone host (route_runoff.cpp) parses the classic format itself: header (dimensions, attributes,
// variables), fixed-size variables and record variables, big-endian, with conversion to double / int.
// netCDF-4/HDF5 files are not supported (convert with `nccopy -k 64-bit-offset`).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace nc3 {

enum Type { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };
inline int type_size(int t) { static const int s[7] = {0, 1, 1, 2, 4, 4, 8}; if (t < 1 || t > 6) throw std::runtime_error("nc3: bad type"); return s[t]; }

struct Attr { int type = 0; std::string text; std::vector<double> values; };
struct Dim { std::string name; uint64_t len = 0; };
struct Var {
    std::string name; std::vector<int> dimids; std::map<std::string, Attr> attrs;
    int type = 0; uint64_t vsize = 0, begin = 0; bool record = false; uint64_t count = 0;   // count = elements per record (or total)
};

inline uint64_t pad4(uint64_t n) { return (n + 3) & ~(uint64_t)3; }

class Reader {
public:
    std::vector<Dim> dims; std::vector<Var> vars; std::map<std::string, Attr> gattrs;
    uint64_t numrecs = 0, recsize = 0; int recdim = -1, version = 1;

    explicit Reader(const std::string &path) : path_(path) {
        f_ = std::fopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("nc3: cannot open " + path);
        unsigned char m[4];
        rd(m, 4);
        if (m[0] != 'C' || m[1] != 'D' || m[2] != 'F' || (m[3] != 1 && m[3] != 2))
            throw std::runtime_error("nc3: " + path + " is not a NetCDF-3 classic/64-bit-offset file (netCDF-4/HDF5 is not supported)");
        version = m[3];
        numrecs = u32();
        // dimensions
        uint32_t tag = u32(), n = u32();
        if (tag != 0 && tag != 0x0A) throw std::runtime_error("nc3: bad dimension list");
        for (uint32_t i = 0; i < n; ++i) { Dim d; d.name = name(); d.len = u32(); if (d.len == 0) recdim = (int)i; dims.push_back(d); }
        read_attrs(gattrs);
        tag = u32(); n = u32();
        if (tag != 0 && tag != 0x0B) throw std::runtime_error("nc3: bad variable list");
        for (uint32_t i = 0; i < n; ++i) {
            Var v; v.name = name();
            const uint32_t nd = u32();
            for (uint32_t k = 0; k < nd; ++k) v.dimids.push_back((int)u32());
            read_attrs(v.attrs);
            v.type = (int)u32(); v.vsize = u32(); v.begin = version == 2 ? u64() : u32();
            v.record = !v.dimids.empty() && v.dimids[0] == recdim;
            v.count = 1;
            for (size_t k = v.record ? 1 : 0; k < v.dimids.size(); ++k) v.count *= dims[v.dimids[k]].len;
            if (v.record) recsize += v.vsize;
            vars.push_back(v);
        }
        int nrec = 0; for (auto &v : vars) nrec += v.record;
        if (nrec == 1) for (auto &v : vars) if (v.record) recsize = v.count * type_size(v.type);   // a lone record variable is not padded
    }
    ~Reader() { if (f_) std::fclose(f_); }
    Reader(const Reader &) = delete;

    const Var *find(const std::string &nm) const { for (auto &v : vars) if (v.name == nm) return &v; return nullptr; }
    const Var &var(const std::string &nm) const { const Var *v = find(nm); if (!v) throw std::runtime_error("nc3: variable '" + nm + "' not found in " + path_); return *v; }
    int dim_index(const std::string &nm) const { for (size_t i = 0; i < dims.size(); ++i) if (dims[i].name == nm) return (int)i; return -1; }
    uint64_t dim_len(const std::string &nm) const {
        const int i = dim_index(nm); if (i < 0) throw std::runtime_error("nc3: dimension '" + nm + "' not found in " + path_);
        return i == recdim ? numrecs : dims[i].len;
    }
    std::string attr_text(const Var &v, const std::string &nm) const { auto it = v.attrs.find(nm); return it == v.attrs.end() ? std::string() : it->second.text; }
    bool attr_value(const Var &v, const std::string &nm, double &out) const {
        auto it = v.attrs.find(nm); if (it == v.attrs.end() || it->second.values.empty()) return false; out = it->second.values[0]; return true;
    }

    // whole fixed-size variable, or records [rec0, rec0+nrec) of a record variable, converted to double
    void read(const Var &v, std::vector<double> &out, uint64_t rec0 = 0, uint64_t nrec = 0) {
        const int ts = type_size(v.type);
        if (!v.record) { out.resize(v.count); raw_.resize(v.count * ts); seek(v.begin); rd(raw_.data(), raw_.size()); convert(v.type, raw_.data(), v.count, out.data()); return; }
        if (rec0 + nrec > numrecs) throw std::runtime_error("nc3: record range beyond the file: " + v.name);
        out.resize(v.count * nrec); raw_.resize(v.count * ts);
        for (uint64_t r = 0; r < nrec; ++r) { seek(v.begin + (rec0 + r) * recsize); rd(raw_.data(), raw_.size()); convert(v.type, raw_.data(), v.count, out.data() + r * v.count); }
    }
    void read_all(const Var &v, std::vector<double> &out) { if (v.record) read(v, out, 0, numrecs); else read(v, out); }
    void read_int(const Var &v, std::vector<int> &out) { std::vector<double> t; read_all(v, t); out.resize(t.size()); for (size_t i = 0; i < t.size(); ++i) out[i] = (int)t[i]; }

private:
    std::string path_; FILE *f_ = nullptr; std::vector<unsigned char> raw_;
    void rd(void *p, size_t n) { if (n && std::fread(p, 1, n, f_) != n) throw std::runtime_error("nc3: unexpected end of " + path_); }
    void seek(uint64_t off) { if (fseeko(f_, (off_t)off, SEEK_SET) != 0) throw std::runtime_error("nc3: seek failed in " + path_); }
    uint32_t u32() { unsigned char b[4]; rd(b, 4); return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
    uint64_t u64() { const uint64_t hi = u32(); return (hi << 32) | u32(); }
    std::string name() { const uint32_t n = u32(); std::string s(pad4(n), '\0'); rd(&s[0], s.size()); s.resize(n); return s; }
    void read_attrs(std::map<std::string, Attr> &out) {
        const uint32_t tag = u32(), n = u32();
        if (tag != 0 && tag != 0x0C) throw std::runtime_error("nc3: bad attribute list");
        for (uint32_t i = 0; i < n; ++i) {
            const std::string nm = name(); Attr a; a.type = (int)u32();
            const uint32_t ne = u32(); const size_t bytes = pad4((uint64_t)ne * type_size(a.type));
            std::vector<unsigned char> buf(bytes); rd(buf.data(), bytes);
            if (a.type == NC_CHAR) { a.text.assign((const char *)buf.data(), ne); while (!a.text.empty() && a.text.back() == '\0') a.text.pop_back(); }
            else { a.values.resize(ne); convert(a.type, buf.data(), ne, a.values.data()); }
            out[nm] = a;
        }
    }
    static void convert(int type, const unsigned char *p, uint64_t n, double *out) {
        for (uint64_t i = 0; i < n; ++i) {
            switch (type) {
                case NC_BYTE: out[i] = (signed char)p[i]; break;
                case NC_CHAR: out[i] = p[i]; break;
                case NC_SHORT: out[i] = (int16_t)(((uint16_t)p[2 * i] << 8) | p[2 * i + 1]); break;
                case NC_INT: { const uint32_t u = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3]; out[i] = (int32_t)u; break; }
                case NC_FLOAT: { uint32_t u = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3]; float f; std::memcpy(&f, &u, 4); out[i] = f; break; }
                case NC_DOUBLE: { uint64_t u = 0; for (int k = 0; k < 8; ++k) u = (u << 8) | p[8 * i + k]; double d; std::memcpy(&d, &u, 8); out[i] = d; break; }
            }
        }
    }
};

// Writer for one file with fixed variables (int / double) and float or double record variables over one unlimited
// dimension; 64-bit offsets (CDF-2).  define_*() first, then end_def(), then put_*().
class Writer {
public:
    explicit Writer(const std::string &path) : path_(path) { f_ = std::fopen(path.c_str(), "wb"); if (!f_) throw std::runtime_error("nc3: cannot create " + path); }
    ~Writer() { close(); }
    int def_dim(const std::string &nm, uint64_t len) { dims_.push_back({nm, len}); return (int)dims_.size() - 1; }   // len 0 = unlimited
    int def_var(const std::string &nm, int type, const std::vector<int> &dimids, const std::map<std::string, std::string> &text_attrs = {}) {
        WVar v; v.name = nm; v.type = type; v.dimids = dimids; v.attrs = text_attrs;
        v.record = !dimids.empty() && dims_[dimids[0]].len == 0;
        v.count = 1; for (size_t k = v.record ? 1 : 0; k < dimids.size(); ++k) v.count *= dims_[dimids[k]].len;
        v.vsize = pad4(v.count * type_size(type));
        vars_.push_back(v); return (int)vars_.size() - 1;
    }
    void global_attr(const std::string &nm, const std::string &val) { gattrs_[nm] = val; }
    void end_def() {
        std::vector<unsigned char> h; build_header(h, 0);
        uint64_t off = pad4(h.size());
        for (auto &v : vars_) if (!v.record) { v.begin = off; off += v.vsize; }
        recsize_ = 0; int nrec = 0;
        for (auto &v : vars_) if (v.record) { v.begin = off + recsize_; recsize_ += v.vsize; ++nrec; }
        if (nrec == 1) for (auto &v : vars_) if (v.record) recsize_ = v.count * type_size(v.type);
        rec0_ = off; header_len_ = pad4(h.size());
        write_header();
    }
    void put_int(int id, const int *data) {
        const WVar &v = vars_[id]; std::vector<unsigned char> b(v.vsize, 0);
        for (uint64_t i = 0; i < v.count; ++i) { const uint32_t u = (uint32_t)data[i]; b[4 * i] = u >> 24; b[4 * i + 1] = u >> 16; b[4 * i + 2] = u >> 8; b[4 * i + 3] = u; }
        seek(v.begin); wr(b.data(), b.size());
    }
    void put_double(int id, const double *data) {
        const WVar &v = vars_[id]; std::vector<unsigned char> b(v.vsize, 0);
        for (uint64_t i = 0; i < v.count; ++i) { uint64_t u; std::memcpy(&u, &data[i], 8); for (int k = 0; k < 8; ++k) b[8 * i + k] = (unsigned char)(u >> (56 - 8 * k)); }
        seek(v.begin); wr(b.data(), b.size());
    }
    // one record of a float (converted from double) or double record variable
    void put_record(int id, uint64_t rec, const double *data) {
        const WVar &v = vars_[id]; std::vector<unsigned char> b(v.count * type_size(v.type));
        for (uint64_t i = 0; i < v.count; ++i) {
            if (v.type == NC_FLOAT) { const float f = (float)data[i]; uint32_t u; std::memcpy(&u, &f, 4); b[4 * i] = u >> 24; b[4 * i + 1] = u >> 16; b[4 * i + 2] = u >> 8; b[4 * i + 3] = u; }
            else { uint64_t u; std::memcpy(&u, &data[i], 8); for (int k = 0; k < 8; ++k) b[8 * i + k] = (unsigned char)(u >> (56 - 8 * k)); }
        }
        seek(v.begin + rec * recsize_); wr(b.data(), b.size());
        if (rec + 1 > numrecs_) numrecs_ = rec + 1;
    }
    void close() {
        if (!f_) return;
        // pad the last record and store numrecs
        fseeko(f_, 0, SEEK_END);
        const uint64_t want = rec0_ + numrecs_ * recsize_; uint64_t have = (uint64_t)ftello(f_);
        while (have < want) { std::fputc(0, f_); ++have; }
        write_header();
        std::fclose(f_); f_ = nullptr;
    }

private:
    struct WVar { std::string name; int type; std::vector<int> dimids; std::map<std::string, std::string> attrs; bool record; uint64_t count, vsize, begin = 0; };
    std::string path_; FILE *f_ = nullptr; std::vector<Dim> dims_; std::vector<WVar> vars_; std::map<std::string, std::string> gattrs_;
    uint64_t recsize_ = 0, rec0_ = 0, numrecs_ = 0, header_len_ = 0;
    void wr(const void *p, size_t n) { if (n && std::fwrite(p, 1, n, f_) != n) throw std::runtime_error("nc3: write failed: " + path_); }
    void seek(uint64_t off) { if (fseeko(f_, (off_t)off, SEEK_SET) != 0) throw std::runtime_error("nc3: seek failed: " + path_); }
    static void p32(std::vector<unsigned char> &h, uint32_t v) { h.push_back(v >> 24); h.push_back(v >> 16); h.push_back(v >> 8); h.push_back(v); }
    static void p64(std::vector<unsigned char> &h, uint64_t v) { p32(h, (uint32_t)(v >> 32)); p32(h, (uint32_t)v); }
    static void pname(std::vector<unsigned char> &h, const std::string &s) { p32(h, (uint32_t)s.size()); for (char c : s) h.push_back((unsigned char)c); while (h.size() % 4) h.push_back(0); }
    static void pattrs(std::vector<unsigned char> &h, const std::map<std::string, std::string> &a) {
        if (a.empty()) { p32(h, 0); p32(h, 0); return; }
        p32(h, 0x0C); p32(h, (uint32_t)a.size());
        for (auto &kv : a) { pname(h, kv.first); p32(h, NC_CHAR); p32(h, (uint32_t)kv.second.size()); for (char c : kv.second) h.push_back((unsigned char)c); while (h.size() % 4) h.push_back(0); }
    }
    void build_header(std::vector<unsigned char> &h, uint64_t numrecs) const {
        h = {'C', 'D', 'F', 2};
        p32(h, (uint32_t)numrecs);
        if (dims_.empty()) { p32(h, 0); p32(h, 0); } else { p32(h, 0x0A); p32(h, (uint32_t)dims_.size()); for (auto &d : dims_) { pname(h, d.name); p32(h, (uint32_t)d.len); } }
        pattrs(h, gattrs_);
        if (vars_.empty()) { p32(h, 0); p32(h, 0); return; }
        p32(h, 0x0B); p32(h, (uint32_t)vars_.size());
        for (auto &v : vars_) {
            pname(h, v.name); p32(h, (uint32_t)v.dimids.size()); for (int d : v.dimids) p32(h, (uint32_t)d);
            pattrs(h, v.attrs); p32(h, (uint32_t)v.type); p32(h, (uint32_t)(v.vsize > 0xffffffffull ? 0xffffffffull : v.vsize)); p64(h, v.begin);
        }
    }
    void write_header() { std::vector<unsigned char> h; build_header(h, numrecs_); h.resize(header_len_, 0); seek(0); wr(h.data(), h.size()); }
};

}  // namespace nc3
