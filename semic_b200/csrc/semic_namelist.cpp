// semic_namelist.cpp -- reader for the &surface_physics namelist group.
//
// Replaces the Fortran namelist READ of surface_physics_par_load (surface_physics.f90:743-780)
// with the same observable semantics: group and variable names are case-insensitive, `!` starts
// a comment, values are separated by blanks, commas or newlines, `d`/`D` exponents, repeat
// counts (`3*""`), array sections (`boundary(2:3) = ...`), consecutive commas are null values,
// foreign groups (&driver, &smb_output) are skipped, keys that do not appear keep the value the
// caller passed in (:752-775), tsticsub = tstic/dble(n_ksub) afterwards (:810).  An unknown key
// inside the group, a missing group or a malformed value is an error (the Fortran run time
// aborts there).
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "semic_b200.h"

namespace semic {

void set_error(const char *fmt, ...);

namespace {

struct Cursor {
    const std::string &s;
    size_t i;
    explicit Cursor(const std::string &str) : s(str), i(0) {}
    bool eof() const { return i >= s.size(); }
    char peek() const { return eof() ? '\0' : s[i]; }
    // skips blanks, newlines and comments; optionally value separators (commas) too
    void skip_ws()
    {
        while (!eof()) {
            const char c = s[i];
            if (c == '!') { while (!eof() && s[i] != '\n') ++i; }
            else if (std::isspace((unsigned char)c)) ++i;
            else break;
        }
    }
};

std::string lower(std::string v)
{
    for (auto &c : v) c = (char)std::tolower((unsigned char)c);
    return v;
}

bool is_ident_start(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
bool is_ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

std::string read_ident(Cursor &c)
{
    std::string id;
    while (!c.eof() && is_ident_char(c.peek())) id.push_back(c.s[c.i++]);
    return lower(id);
}

// Looks ahead (without consuming) for `name [ ( ... ) ] =`
bool at_assignment(const Cursor &c0)
{
    Cursor c(c0.s);
    c.i = c0.i;
    if (!is_ident_start(c.peek())) return false;
    while (!c.eof() && (is_ident_char(c.peek()) || c.peek() == '%')) ++c.i;
    c.skip_ws();
    if (c.peek() == '(') {
        while (!c.eof() && c.peek() != ')') ++c.i;
        if (c.eof()) return false;
        ++c.i;
        c.skip_ws();
    }
    return c.peek() == '=';
}

struct Value {
    bool null = false;
    bool is_string = false;
    std::string text;
};

// one value token (without repeat handling); returns false on syntax error
bool read_scalar(Cursor &c, Value &v)
{
    const char q = c.peek();
    if (q == '"' || q == '\'') {
        ++c.i;
        v.is_string = true;
        for (;;) {
            if (c.eof()) return false;
            const char ch = c.s[c.i++];
            if (ch == q) {
                if (c.peek() == q) { v.text.push_back(q); ++c.i; }   // doubled quote
                else break;
            } else v.text.push_back(ch);
        }
        return true;
    }
    while (!c.eof()) {
        const char ch = c.peek();
        if (std::isspace((unsigned char)ch) || ch == ',' || ch == '/' || ch == '!' || ch == '*') break;
        v.text.push_back(ch);
        ++c.i;
    }
    return !v.text.empty();
}

bool to_double(const std::string &t, double &out)
{
    std::string u = t;
    for (auto &ch : u) if (ch == 'd' || ch == 'D') ch = 'e';
    // Fortran also accepts "1.5+3" (exponent without letter); not used by any shipped namelist
    errno = 0;
    char *end = nullptr;
    out = std::strtod(u.c_str(), &end);
    if (end == u.c_str() || *end != '\0') return false;
    // ERANGE on underflow (a subnormal or zero result) is a value Fortran's list-directed input accepts; overflow is an error
    return errno == 0 || (errno == ERANGE && std::fabs(out) < 1.0);
}

bool to_int(const std::string &t, int &out)
{
    errno = 0;
    char *end = nullptr;
    const long v = std::strtol(t.c_str(), &end, 10);
    if (end == t.c_str() || *end != '\0' || errno != 0) return false;
    out = (int)v;
    return true;
}

void put_fstring(char *dst, const std::string &v)
{
    std::memset(dst, ' ', SEMIC_STRLEN);                         // blank padded like character(len=256)
    std::memcpy(dst, v.data(), v.size() < (size_t)SEMIC_STRLEN ? v.size() : (size_t)SEMIC_STRLEN);
}

struct Key {
    const char *name;
    int kind;            // 0 double, 1 int, 2 string, 3 string array
    size_t offset;
};

#define KD(n) {#n, 0, offsetof(semic_par, n)}
const Key kKeys[] = {
    {"boundary", 3, offsetof(semic_par, boundary)},
    KD(tstic), KD(ceff), KD(csh), KD(clh), KD(alb_smax), KD(alb_smin), KD(albi), KD(albl), KD(tmin), KD(tmax),
    KD(hcrit), KD(rcrit), KD(amp), KD(tau_a), KD(tau_f), KD(w_crit), KD(mcrit), KD(afac), KD(tmid),
    {"alb_scheme", 2, offsetof(semic_par, alb_scheme)},
    {"n_ksub", 1, offsetof(semic_par, n_ksub)},
};
#undef KD

}  // namespace

int namelist_load(semic_par *par, const char *filename)
{
    FILE *fp = std::fopen(filename, "rb");
    if (!fp) {
        set_error("surface_physics_par_load: cannot open '%s'", filename);
        return SEMIC_ERR_IO;
    }
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
    std::fclose(fp);

    Cursor c(text);
    // locate "&surface_physics" (or "$surface_physics"); groups in between are skipped wholesale
    bool found = false;
    while (!c.eof()) {
        c.skip_ws();
        if (c.eof()) break;
        const char ch = c.peek();
        if (ch == '&' || ch == '$') {
            ++c.i;
            const std::string g = read_ident(c);
            if (g == "surface_physics") { found = true; break; }
            // skip this foreign group up to its terminating '/' (outside strings/comments)
            while (!c.eof()) {
                const char k = c.peek();
                if (k == '!') { while (!c.eof() && c.peek() != '\n') ++c.i; }
                else if (k == '"' || k == '\'') {
                    ++c.i;
                    while (!c.eof() && c.peek() != k) ++c.i;
                    if (!c.eof()) ++c.i;
                } else if (k == '/') { ++c.i; break; }
                else if ((k == '&' || k == '$') && lower(text.substr(c.i + 1, 3)) == "end") { c.i += 4; break; }
                else ++c.i;
            }
        } else {
            while (!c.eof() && c.peek() != '\n') ++c.i;      // junk outside groups is ignored
        }
    }
    if (!found) {
        set_error("surface_physics_par_load: group &surface_physics not found in '%s'", filename);
        return SEMIC_ERR_PARSE;
    }

    for (;;) {
        c.skip_ws();
        while (c.peek() == ',') { ++c.i; c.skip_ws(); }
        if (c.eof()) { set_error("namelist '%s': end of file inside &surface_physics", filename); return SEMIC_ERR_PARSE; }
        if (c.peek() == '/') break;
        if ((c.peek() == '&' || c.peek() == '$') && lower(text.substr(c.i + 1, 3)) == "end") break;
        if (!is_ident_start(c.peek())) {
            set_error("namelist '%s': unexpected character '%c' at offset %zu", filename, c.peek(), c.i);
            return SEMIC_ERR_PARSE;
        }
        const std::string name = read_ident(c);
        const Key *key = nullptr;
        for (const Key &k : kKeys) if (name == k.name) { key = &k; break; }
        if (!key) {
            set_error("namelist '%s': '%s' is not a member of &surface_physics", filename, name.c_str());
            return SEMIC_ERR_PARSE;
        }
        c.skip_ws();
        int first = 1, last = -1;                                  // 1-based section of an array key
        if (c.peek() == '(') {
            ++c.i;
            std::string spec;
            while (!c.eof() && c.peek() != ')') spec.push_back(c.s[c.i++]);
            if (c.eof()) { set_error("namelist '%s': unterminated subscript of '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
            ++c.i;
            const size_t colon = spec.find(':');
            std::string a = colon == std::string::npos ? spec : spec.substr(0, colon);
            std::string b = colon == std::string::npos ? spec : spec.substr(colon + 1);
            auto strip = [](std::string &s) {
                size_t p = 0; while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
                size_t q = s.size(); while (q > p && std::isspace((unsigned char)s[q - 1])) --q;
                s = s.substr(p, q - p);
            };
            strip(a); strip(b);
            if (key->kind != 3 || (!a.empty() && !to_int(a, first)) || (!b.empty() && !to_int(b, last)) ||
                first < 1 || first > SEMIC_NBOUNDARY) {
                set_error("namelist '%s': bad subscript of '%s'", filename, name.c_str());
                return SEMIC_ERR_PARSE;
            }
            if (colon == std::string::npos) last = first;
            c.skip_ws();
        }
        if (c.peek() != '=') { set_error("namelist '%s': '=' expected after '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
        ++c.i;

        // collect the value list
        std::vector<Value> vals;
        bool pending_sep = true;                                   // a value may follow
        for (;;) {
            c.skip_ws();
            if (c.eof() || c.peek() == '/' || at_assignment(c)) break;
            if ((c.peek() == '&' || c.peek() == '$')) break;
            if (c.peek() == ',') {
                if (pending_sep) { Value nv; nv.null = true; vals.push_back(nv); }   // null value
                pending_sep = true;
                ++c.i;
                continue;
            }
            Value v;
            if (!read_scalar(c, v)) { set_error("namelist '%s': bad value for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
            int repeat = 1;
            if (!v.is_string && c.peek() == '*') {                 // r*value
                if (!to_int(v.text, repeat) || repeat < 1) { set_error("namelist '%s': bad repeat count for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
                ++c.i;
                Value w;
                const char nx = c.peek();
                if (c.eof() || std::isspace((unsigned char)nx) || nx == ',' || nx == '/') w.null = true;
                else if (!read_scalar(c, w)) { set_error("namelist '%s': bad value for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
                v = w;
            }
            for (int r = 0; r < repeat; ++r) vals.push_back(v);
            pending_sep = false;
        }

        char *base = reinterpret_cast<char *>(par) + key->offset;
        if (key->kind == 3) {
            int idx = first;
            for (const Value &v : vals) {
                if (idx > SEMIC_NBOUNDARY || (last >= 1 && idx > last)) { set_error("namelist '%s': too many values for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
                if (!v.null) {
                    if (!v.is_string) { set_error("namelist '%s': string expected for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
                    put_fstring(base + (size_t)(idx - 1) * SEMIC_STRLEN, v.text);
                }
                ++idx;
            }
        } else {
            if (vals.size() > 1) { set_error("namelist '%s': too many values for scalar '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
            if (vals.size() == 1 && !vals[0].null) {
                const Value &v = vals[0];
                if (key->kind == 2) {
                    if (!v.is_string) { set_error("namelist '%s': string expected for '%s'", filename, name.c_str()); return SEMIC_ERR_PARSE; }
                    put_fstring(base, v.text);
                } else if (key->kind == 0) {
                    double d;
                    if (v.is_string || !to_double(v.text, d)) { set_error("namelist '%s': bad real '%s' for '%s'", filename, v.text.c_str(), name.c_str()); return SEMIC_ERR_PARSE; }
                    *reinterpret_cast<double *>(base) = d;
                } else {
                    int k;
                    if (v.is_string || !to_int(v.text, k)) { set_error("namelist '%s': bad integer '%s' for '%s'", filename, v.text.c_str(), name.c_str()); return SEMIC_ERR_PARSE; }
                    *reinterpret_cast<int *>(base) = k;
                }
            }
        }
    }
    par->tsticsub = par->tstic / (double)par->n_ksub;              // surface_physics.f90:810
    return SEMIC_OK;
}

}  // namespace semic
