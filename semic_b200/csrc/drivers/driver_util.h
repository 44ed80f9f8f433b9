// driver_util.h -- host-side helpers of the two command-line drivers (example.x, run_particles.x): the &driver
// namelist group (example/example.f90:26, optimize/run_particles.f90:30), the list-directed ASCII readers of
// utils.f90:44-95, and error handling over the C-ABI.  Host code only; everything that computes is in libsemic_b200.so.
#pragma once

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "semic_b200.h"

namespace drv {

#define GREEN "\x1B[32m"
#define RED   "\x1B[31m"
#define ULINE "\x1B[4;37m"
#define RESET "\x1B[0m"

[[noreturn]] inline void die(const std::string &msg)
{
    std::fprintf(stderr, " " RED "Error:" RESET " %s\n", msg.c_str());
    std::exit(1);
}
inline void check(int rc, const char *what)
{
    if (rc != SEMIC_OK) die(std::string(what) + ": " + semic_last_error());
}

// namelist /driver/ nloop, ntime, nx, loi_mask, input_forcing, output, validation
struct DriverNml {
    int nloop = 0, ntime = 0, nx = 0;
    std::vector<int> loi_mask;      // loi_mask(100) of the reference; longer here (entries past 100 keep the default) so that nx > 100 runs
    std::string input_forcing, output, validation;
};

inline std::string lower(std::string s)
{
    for (char &c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

// Reads the &driver group: `name = value`, `name(lo:hi) = v1,v2,...`, quoted strings, `!` comments, `,` or blanks
// between items, `d` exponents, the group ends at `/`.  Unknown names are an error, as in a Fortran namelist read.
inline DriverNml read_driver_namelist(const std::string &path, int mask_default)
{
    std::ifstream f(path);
    if (!f) die("file " ULINE + path + RESET " not found.");
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str(), body;
    {   // strip comments outside strings
        std::string t;
        char q = 0;
        for (size_t i = 0; i < text.size(); ++i) {
            const char c = text[i];
            if (q) { if (c == q) q = 0; t += c; continue; }
            if (c == '"' || c == '\'') { q = c; t += c; continue; }
            if (c == '!') { while (i < text.size() && text[i] != '\n') ++i; t += '\n'; continue; }
            t += c;
        }
        text = t;
    }
    const std::string low = lower(text);
    size_t g = low.find("&driver");
    if (g == std::string::npos) die("namelist group &driver not found in " + path);
    g += 7;
    size_t e = g;
    {
        char q = 0;
        for (; e < text.size(); ++e) {
            const char c = text[e];
            if (q) { if (c == q) q = 0; continue; }
            if (c == '"' || c == '\'') { q = c; continue; }
            if (c == '/') break;
        }
    }
    body = text.substr(g, e - g);
    DriverNml d;
    d.loi_mask.assign(100, mask_default);                           // example.f90:51, run_particles.f90:61
    size_t i = 0;
    auto skip = [&]() { while (i < body.size() && (std::isspace((unsigned char)body[i]) || body[i] == ',')) ++i; };
    while (true) {
        skip();
        if (i >= body.size()) break;
        size_t j = i;
        while (j < body.size() && (std::isalnum((unsigned char)body[j]) || body[j] == '_')) ++j;
        if (j == i) die("namelist &driver: unexpected character '" + std::string(1, body[i]) + "'");
        const std::string name = lower(body.substr(i, j - i));
        i = j;
        int lo = 1;
        while (i < body.size() && std::isspace((unsigned char)body[i])) ++i;
        if (i < body.size() && body[i] == '(') {                     // array section (lo:hi) or element (k)
            const size_t c = body.find(')', i);
            if (c == std::string::npos) die("namelist &driver: unbalanced '('");
            const std::string sec = body.substr(i + 1, c - i - 1);
            if (!sec.empty() && sec[0] != ':') lo = std::atoi(sec.c_str());
            i = c + 1;
        }
        while (i < body.size() && std::isspace((unsigned char)body[i])) ++i;
        if (i >= body.size() || body[i] != '=') die("namelist &driver: '=' expected after " + name);
        ++i;
        // values until the next `name =` / `name(` `=` or the end of the group
        std::vector<std::string> vals;
        while (true) {
            skip();
            if (i >= body.size()) break;
            if (body[i] == '"' || body[i] == '\'') {
                const char q = body[i];
                const size_t c = body.find(q, i + 1);
                if (c == std::string::npos) die("namelist &driver: unterminated string");
                vals.push_back(body.substr(i + 1, c - i - 1));
                i = c + 1;
                continue;
            }
            size_t k = i;
            while (k < body.size() && !std::isspace((unsigned char)body[k]) && body[k] != ',' && body[k] != '=' && body[k] != '(') ++k;
            size_t look = k;                                         // is this token the next assignment's name?
            while (look < body.size() && std::isspace((unsigned char)body[look])) ++look;
            if (look < body.size() && (body[look] == '=' || body[look] == '(') &&
                (std::isalpha((unsigned char)body[i]) || body[i] == '_')) break;
            vals.push_back(body.substr(i, k - i));
            i = k;
        }
        auto num = [&](const std::string &t) {
            std::string u = t;
            for (char &c : u) if (c == 'd' || c == 'D') c = 'e';
            return std::atof(u.c_str());
        };
        if (name == "nloop" || name == "ntime" || name == "nx") {
            if (vals.empty()) die("namelist &driver: no value for " + name);
            const int v = (int)num(vals[0]);
            (name == "nloop" ? d.nloop : name == "ntime" ? d.ntime : d.nx) = v;
        } else if (name == "input_forcing" || name == "output" || name == "validation") {
            if (vals.empty()) die("namelist &driver: no value for " + name);
            (name == "input_forcing" ? d.input_forcing : name == "output" ? d.output : d.validation) = vals[0];
        } else if (name == "loi_mask") {
            for (size_t k = 0; k < vals.size(); ++k) {
                const int idx = lo - 1 + (int)k;
                if (idx < 0 || idx >= 100) die("namelist &driver: loi_mask index out of range (loi_mask(100))");
                d.loi_mask[idx] = (int)num(vals[k]);
            }
        } else {
            die("namelist &driver: unknown variable " + name);
        }
    }
    // the reference stops at nx = 100 (loi_mask(100), example.f90:23, Q13); larger grids run here with the default mask
    // (2 in example.f90:51 and run_particles.f90:61) beyond point 100 -- an extension, the first 100 points behave alike
    if (d.nx < 1) die("namelist &driver: nx must be positive");
    if ((size_t)d.nx > d.loi_mask.size()) d.loi_mask.resize((size_t)d.nx, mask_default);
    if (d.ntime < 1 || d.nloop < 1) die("namelist &driver: ntime and nloop must be positive");
    return d;
}

// read_forcing / read_validation (utils.f90:44-95): ntime records of nvar*nx list-directed reals, variable-major
// within a record.  Returns [day][var][nx].
inline std::vector<double> read_table(const std::string &path, int ntime, int nvar, int nx, const char *prog, const char *what)
{
    std::printf(" " GREEN "%s:" RESET " Read %s from file " ULINE "%s" RESET "\n", prog, what, path.c_str());
    std::FILE *f = std::fopen(path.c_str(), "r");
    if (!f) die("cannot open " + path);
    std::vector<double> a((size_t)ntime * nvar * nx);
    std::string line;
    std::vector<char> buf(1 << 16);
    for (int d = 0; d < ntime; ++d) {
        int got = 0;
        const int need = nvar * nx;
        while (got < need) {                                          // a record may continue over several lines
            line.clear();
            bool eol = false;
            while (!eol && std::fgets(buf.data(), (int)buf.size(), f)) {
                line += buf.data();
                eol = !line.empty() && line.back() == '\n';
            }
            if (line.empty()) die(std::string("end of file in ") + path + " at record " + std::to_string(d + 1));
            char *p = &line[0];
            while (got < need) {
                while (*p && (std::isspace((unsigned char)*p) || *p == ',')) ++p;
                if (!*p) break;
                char *q = p;
                while (*q && !std::isspace((unsigned char)*q) && *q != ',') { if (*q == 'd' || *q == 'D') *q = 'e'; ++q; }
                char *end = nullptr;
                const double v = std::strtod(p, &end);
                if (end == p) die(std::string("bad number in ") + path + " at record " + std::to_string(d + 1));
                a[(size_t)d * need + got++] = v;
                p = q;
            }
        }
    }
    std::fclose(f);
    return a;
}

// semic_par with defined contents before the first par_load (the reference's incoming `par` is whatever the program
// holds, Q10): zeros, blank strings
inline void blank_par(semic_par &par)
{
    std::memset(&par, 0, sizeof par);
    std::memset(par.name, ' ', sizeof par.name);
    std::memset(par.boundary, ' ', sizeof par.boundary);
    std::memset(par.alb_scheme, ' ', sizeof par.alb_scheme);
}

}  // namespace drv
