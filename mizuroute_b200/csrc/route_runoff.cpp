// route_runoff -- stand-alone host of the B200 routing library, the counterpart of the reference's
// PROGRAM route_runoff (route/build/src/standalone/route_runoff.f90:5-117):
//
//     route_runoff <control file> [--batch N] [--device-ingest] [--device-history] [--dry-run [--dump-forcing FILE] [--dump-remap FILE]]
//
//   init_model      read_control (read_control.f90:18: lines "<key> value ! comment", '!' comment lines, unknown key =
//                   error) and the parameter namelist &HSLOPE/&IRF_UH/&KWT (read_param.f90:12)
//   init_data       river network netCDF (read_streamSeg.f90:44: seg/hru dimensions, variable names from the
//                   <varname_*> keys) -> mr_set_network;  runoff netCDF(s) listed by <fname_qsim> (model_setup.f90
//                   inFile_pop): time axis, HRU ids; cold start (init_model_data.f90:399-463)
//   time loop       get_hru_runoff (get_basin_runoff.f90:19: record read + sort_flux, process_remap.f90:271: forcing
//                   HRU id -> network HRU index, HRUs without forcing and negative values -> 0) -> mr_step_batch
//                   (mpi_route/main_route) -> history output (write_simoutput_pio.f90:140: float32 [time, seg])
//
// Everything numerical happens behind the C ABI (include/mizuroute_b200.h); this file is I/O and bookkeeping.
// The reference's Fortran host cannot be built in this image (no Fortran compiler, no netCDF/PIO); NetCDF-3
// classic / 64-bit-offset files are read and written with nc3.h.  Restrictions (each one is an explicit error):
// <outputFrequency> = n steps or daily; standard / proleptic_gregorian / noleap calendars.  <restart_write> never | last |
// specified | yearly | monthly | daily.
// <is_remap> T: polygon [time, hru] or gridded [time, lat, lon] forcing, remapped on the device; any ratio of <dt_qsim> to
// the forcing interval; <newFileFrequency> single | daily | monthly | yearly.  <is_flux_wm> / <is_vol_wm> T: one water-management
// netCDF <fname_wm> with [time, seg] variables.
// <qmodOption> 1: gauge metadata csv <gageMetaFile> + gauge netCDF <fname_gageObs> (in <ancil_dir>) -> mr_set_da / mr_upload_obs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mizuroute_b200.h"
#include "nc3.h"

namespace {

[[noreturn]] void die(int ierr, const std::string &msg) {       // handle_err, model_utils.f90:45-57
    std::fprintf(stderr, "FATAL ERROR (ierr=%d): %s\n", ierr, msg.c_str());
    std::exit(ierr ? ierr : 1);
}

std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string lower(std::string s) { for (auto &c : s) c = (char)std::tolower((unsigned char)c); return s; }

// ---- control file -----------------------------------------------------------------------------------------
const char *KNOWN_KEYS[] = {
    "ancil_dir", "input_dir", "output_dir", "restart_dir", "case_name", "sim_start", "sim_end", "continue_run", "route_opt", "doesBasinRoute",
    "dt_qsim", "floodplain", "hw_drain_point", "tracer", "is_lake_sim", "lakeRegulate", "LakeInputOption", "is_flux_wm", "is_vol_wm",
    "is_vol_wm_jumpstart", "scale_factor_runoff", "offset_value_runoff", "scale_factor_Ep", "offset_value_Ep", "is_Ep_upward_negative",
    "scale_factor_prec", "offset_value_prec", "min_length_route", "ntopAugmentMode", "units_qsim", "units_cc", "dt_ro", "input_fillvalue",
    "ro_calendar", "ro_time_units", "ro_time_stamp", "runoffMin", "dt_wm", "is_remap", "restart_write", "restart_date", "restart_month",
    "restart_day", "restart_hour", "param_nml", "qmodOption", "qBlendPeriod", "QerrTrend", "hydGeometryOption", "topoNetworkOption",
    "computeReachList", "gageMetaFile", "outputAtGage", "strlen_gageSite", "fname_gageObs", "vname_gageFlow", "vname_gageSite", "vname_gageTime",
    "dname_gageSite", "dname_gageTime", "pio_netcdf_format", "pio_netcdf_type", "debug", "seg_outlet",
    "desireId", "checkMassBalance", "maxPfafLen", "pfafMissing", "time_units", "newFileFrequency", "outputFrequency", "outputNameOption",
    "histTimeStamp_offset", "basRunoff", "instRunoff", "dlayRunoff", "sumUpstreamRunoff", "KWTroutedRunoff", "IRFroutedRunoff",
    "KWroutedRunoff", "DWroutedRunoff", "MCroutedRunoff", "IRFvolume", "KWTvolume", "KWvolume", "MCvolume", "DWvolume", "KWfloodVolume",
    "MCfloodVolume", "DWfloodVolume", "KWheight", "MCheight", "DWheight", "localSolute", "soluteFlux", "soluteMass", "outputInflow",
    "KWTinflow", "IRFinflow", "KWinflow", "MCinflow", "DWinflow", "qgwl_runoff_option", "bypass_routing_option", "correct_area", "ice_runoff"};

struct Control {
    std::map<std::string, std::string> kv;
    bool has(const std::string &k) const { return kv.count(k) != 0; }
    std::string str(const std::string &k, const std::string &dflt) const { auto it = kv.find(k); return it == kv.end() ? dflt : it->second; }
    std::string need(const std::string &k) const { auto it = kv.find(k); if (it == kv.end()) die(20, "read_control/<" + k + "> must be given"); return it->second; }
    double num(const std::string &k, double dflt) const { auto it = kv.find(k); if (it == kv.end()) return dflt; char *e; const double v = std::strtod(it->second.c_str(), &e); if (e == it->second.c_str()) die(20, "read_control/cannot read a number from <" + k + ">"); return v; }
    bool flag(const std::string &k, bool dflt) const { auto it = kv.find(k); if (it == kv.end()) return dflt; const std::string v = lower(it->second); return v == "t" || v == ".true." || v == "true"; }
};

Control read_control(const std::string &path) {                 // read_control.f90:18-116
    std::ifstream in(path);
    if (!in) die(20, "read_control/cannot open control file " + path);
    std::set<std::string> known(std::begin(KNOWN_KEYS), std::end(KNOWN_KEYS));
    Control c; std::string line; int lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        if (line.find('\t') != std::string::npos) die(20, "read_control/TAB in control file, line " + std::to_string(lineno));
        std::string s = trim(line);
        if (s.empty() || s[0] == '!') continue;
        if (s[0] != '<') die(20, "read_control/expect '<' at the start of line " + std::to_string(lineno));
        const size_t gt = s.find('>');
        if (gt == std::string::npos) die(20, "read_control/missing '>' in line " + std::to_string(lineno));
        const std::string key = s.substr(1, gt - 1);
        std::string val = s.substr(gt + 1);
        const size_t bang = val.find('!');
        if (bang == std::string::npos) die(20, "read_control/missing '!' delimiter in line " + std::to_string(lineno));   // read_control.f90:105-109
        val = trim(val.substr(0, bang));
        const bool nameKey = key.rfind("varname_", 0) == 0 || key.rfind("vname_", 0) == 0 || key.rfind("dname_", 0) == 0 || key.rfind("fname_", 0) == 0;
        if (!nameKey && !known.count(key)) die(81, "read_control/unknown control key <" + key + ">");       // read_control.f90:374-377
        c.kv[key] = val;
    }
    return c;
}

// &HSLOPE fshape,tscale / &IRF_UH velo,diff / &KWT mann_n,wscale /   (read_param.f90:26-38)
void read_param_nml(const std::string &path, mr_options &o) {
    std::ifstream in(path);
    if (!in) die(20, "read_param/cannot open namelist " + path);
    std::stringstream ss; ss << in.rdbuf();
    std::string t = ss.str();
    // strip comments
    std::string clean; bool com = false;
    for (char ch : t) { if (ch == '!') com = true; if (ch == '\n') com = false; if (!com) clean += ch; }
    auto get = [&](const std::string &name, double &dst) {
        const std::string lc = lower(clean);
        size_t p = 0;
        while ((p = lc.find(name, p)) != std::string::npos) {
            const bool left = p == 0 || !(std::isalnum((unsigned char)lc[p - 1]) || lc[p - 1] == '_');
            size_t q = p + name.size();
            while (q < lc.size() && std::isspace((unsigned char)lc[q])) ++q;
            if (left && q < lc.size() && lc[q] == '=') {
                std::string v = clean.substr(q + 1, 64);
                for (auto &ch : v) if (ch == 'd' || ch == 'D') ch = 'e';       // Fortran 1.0d0
                dst = std::strtod(v.c_str(), nullptr);
                return;
            }
            p = q;
        }
    };
    get("fshape", o.fshape); get("tscale", o.tscale); get("velo", o.velo); get("diff", o.diff); get("mann_n", o.mann_n); get("wscale", o.wscale);
}

// ---- time ---------------------------------------------------------------------------------------------------
// days since 1970-01-01 in the proleptic Gregorian calendar / in a 365-day calendar
long long days_from_civil(long long y, int m, int d, bool noleap) {
    if (noleap) { static const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334}; return (y - 1970) * 365 + cum[m - 1] + (d - 1); }
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400), doy = (153u * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1, doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
}
double parse_datetime(const std::string &s, bool noleap) {      // "yyyy-mm-dd [hh:mm:ss]" -> seconds since 1970-01-01
    int y = 0, mo = 1, d = 1, h = 0, mi = 0; double sec = 0.0;
    const int n = std::sscanf(s.c_str(), "%d-%d-%d %d:%d:%lf", &y, &mo, &d, &h, &mi, &sec);
    if (n < 3) { if (std::sscanf(s.c_str(), "%d-%d-%dT%d:%d:%lf", &y, &mo, &d, &h, &mi, &sec) < 3) die(20, "datetime/cannot parse date '" + s + "'"); }
    return (double)days_from_civil(y, mo, d, noleap) * 86400.0 + h * 3600.0 + mi * 60.0 + sec;
}
// "<unit> since <date>" -> seconds per unit and reference epoch (model_setup.f90 inFile_pop: t_unit / convTime2sec)
void parse_time_units(const std::string &units, bool noleap, double &scale, double &epoch) {
    const std::string u = lower(trim(units));
    const size_t p = u.find("since");
    if (p == std::string::npos) die(20, "inFile_pop/time units must be '<unit> since yyyy-mm-dd hh:mm:ss': " + units);
    const std::string unit = trim(u.substr(0, p));
    if (unit.rfind("sec", 0) == 0) scale = 1.0; else if (unit.rfind("min", 0) == 0) scale = 60.0; else if (unit.rfind("hour", 0) == 0 || unit == "h" || unit == "hr") scale = 3600.0;
    else if (unit.rfind("day", 0) == 0) scale = 86400.0; else die(20, "inFile_pop/<time_units>= " + unit + ": must be seconds, minutes, hours or days");
    epoch = parse_datetime(trim(units.substr(p + 5)), noleap);
}

// civil date of seconds since 1970-01-01 (proleptic Gregorian; noleap: 365-day years)
struct Civil { int y, mo, d, sod; };
Civil civil_from_sec(double t, bool noleap) {
    const long long days = (long long)std::floor((t + 1e-6) / 86400.0);
    Civil cv; cv.sod = (int)std::llround(t - (double)days * 86400.0);
    if (noleap) { long long yy = days >= 0 ? days / 365 : -((-days + 364) / 365); cv.y = 1970 + (int)yy; int doy = (int)(days - yy * 365);
                  static const int ml[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31}; int m = 0; while (doy >= ml[m]) doy -= ml[m++]; cv.d = doy + 1; cv.mo = m + 1; }
    else { const long long z = days + 719468, era = (z >= 0 ? z : z - 146096) / 146097; const unsigned doe = (unsigned)(z - era * 146097), yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
           const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100), mp = (5 * doy + 2) / 153; cv.d = (int)(doy - (153 * mp + 2) / 5 + 1); cv.mo = (int)(mp < 10 ? mp + 3 : mp - 9); cv.y = (int)(yoe + era * 400 + (cv.mo <= 2)); }
    return cv;
}
// history files: a new file when the year / month / day of the step's start changes (newFileAlarm, write_simoutput_pio.f90:110-135),
// named <case>.h.<yyyy-mm-dd-sssss | yyyy-mm | yyyy>.nc after the first step it holds (get_hfilename, :329-392)
bool new_file_alarm(const std::string &freq, const Civil &prev, const Civil &cur) {
    if (freq == "yearly") return cur.y != prev.y;
    if (freq == "monthly") return cur.mo != prev.mo;
    if (freq == "daily") return cur.d != prev.d;
    return false;                                                // single: only the first step opens a file
}
std::string hist_stamp(const std::string &freq, const Civil &cv) {
    char b[64];
    if (freq == "monthly") std::snprintf(b, sizeof b, "%04d-%02d", cv.y, cv.mo);
    else if (freq == "yearly") std::snprintf(b, sizeof b, "%04d", cv.y);
    else std::snprintf(b, sizeof b, "%04d-%02d-%02d-%05d", cv.y, cv.mo, cv.d, cv.sod);
    return b;
}

std::string join_path(const std::string &dir, const std::string &f) { return (!f.empty() && f[0] == '/') ? f : dir + f; }

struct Forcing { std::string path; size_t nTime = 0; std::vector<double> tsec; };     // record times, seconds since 1970

// ---- restart files in the reference's schema (write_restart_pio.f90:544,983-1134; read_restart.f90:402-470) ---------
// Fortran dimensions (seg, tdh|wave) are (tdh|wave, seg) in file order.  wave = MR_KW_SLOTS because a reach can hold
// MAXQPAR+1 waves between steps.
void check(int ierr, const char *msg) { if (ierr) die(ierr, msg); }

// The open output period of the history file, as the reference carries it through a restart (write_restart_pio.f90:315-324,
// 1324-1480; histVars%read_restart): `nt` steps accumulated so far, `history_time` = start of the period and end of the last
// step in it, and the running SUMS of every history variable under its history-file name ([seg], basRunoff [hru]).
struct HistVar { std::string name; std::vector<double> *data; size_t offset, n; bool hru; };
struct HistState { int nt = 0; double tb[2] = {0.0, 0.0}; std::vector<HistVar> vars; };

void write_restart(mr_handle h, const std::string &path, const mr_options &o, const std::vector<int> &segId, double T0, long steps, bool da = false,
                   const HistState *hist = nullptr) {
    char msg[MR_STRLEN];
    const size_t N = segId.size();
    const int nb = (int)mr_get_info(h, MR_INFO_NTDH_BAS), mx = (int)mr_get_info(h, MR_INFO_MAXTDH), W = MR_KW_SLOTS;
    bool irf = false, kwt = false;
    for (int r = 0; r < o.n_routes; ++r) { irf |= o.route_methods[r] == MR_IMPULSE_RESPONSE_FUNC; kwt |= o.route_methods[r] == MR_KINEMATIC_WAVE_TRACKING; }
    nc3::Writer w(path);
    const int dSeg = w.def_dim("seg", N), dTdh = w.def_dim("tdh", nb), dIrf = w.def_dim("tdh_irf", mx), dWave = w.def_dim("wave", W), dTb = w.def_dim("tbound", 2);
    const int vId = w.def_var("reachID", nc3::NC_INT, {dSeg}), vTb = w.def_var("time_bound", nc3::NC_DOUBLE, {dTb}, {{"units", "sec"}, {"long_name", "time bound at last time step"}});      // TSEC(1:2) of the step just completed (write_restart_pio.f90:327,812; stored in double here, ncd_float there)
    const int vBq = w.def_var("basin_q", nc3::NC_DOUBLE, {dSeg}, {{"units", "m3/s"}}), vQf = w.def_var("qfuture", nc3::NC_DOUBLE, {dTdh, dSeg}, {{"units", "m3/s"}});
    int vNqf = -1, vIq = -1, vIv = -1, vNw = -1, vTe = -1, vTx = -1, vQw = -1, vQm = -1, vRt = -1, vKv = -1;
    if (irf) { vNqf = w.def_var("numQF", nc3::NC_INT, {dSeg}); vIq = w.def_var("irf_qfuture", nc3::NC_DOUBLE, {dIrf, dSeg}, {{"units", "m3/s"}}); vIv = w.def_var("volume_irf", nc3::NC_DOUBLE, {dSeg}, {{"units", "m3"}}); }
    if (kwt) { vNw = w.def_var("numWaves", nc3::NC_INT, {dSeg}); vTe = w.def_var("tentry", nc3::NC_DOUBLE, {dWave, dSeg}, {{"units", "s"}}); vTx = w.def_var("texit", nc3::NC_DOUBLE, {dWave, dSeg}, {{"units", "s"}});
               vQw = w.def_var("qwave", nc3::NC_DOUBLE, {dWave, dSeg}, {{"units", "m2/s"}}); vQm = w.def_var("qwave_mod", nc3::NC_DOUBLE, {dWave, dSeg}, {{"units", "m2/s"}});
               vRt = w.def_var("routed", nc3::NC_INT, {dWave, dSeg}); vKv = w.def_var("volume_kwt", nc3::NC_DOUBLE, {dSeg}, {{"units", "m3"}}); }
    static const char *eul[3] = {"kw", "mc", "dw"}; static const int molN[3] = {20, 2, 20};      // popMetadat.f90:283-295
    int vMol[3] = {-1, -1, -1}, vMolVol[3] = {-1, -1, -1};
    for (int q = 0; q < o.n_routes; ++q) {
        const int m = o.route_methods[q];
        if (m < MR_KINEMATIC_WAVE) continue;
        const std::string e = eul[m - 3];
        const int dMol = w.def_dim("mol_" + e, molN[m - 3]);
        vMol[m - 3] = w.def_var("q_sub_" + e, nc3::NC_DOUBLE, {dMol, dSeg}, {{"units", "m3/s"}});
        vMolVol[m - 3] = w.def_var("volume_" + e, nc3::NC_DOUBLE, {dSeg}, {{"units", "m3"}});
    }
    // discharge error of the corrected methods: written only under data assimilation (write_restart_pio.f90:1021-1030 and siblings)
    static const char *qerrName[6] = {nullptr, "qerror_irf", nullptr, "qerror_kw", "qerror_mc", "qerror_dw"};      // popMetadat.f90:285-300
    int vQerr[6] = {-1, -1, -1, -1, -1, -1};
    if (da) for (int q = 0; q < o.n_routes; ++q) { const int m = o.route_methods[q]; if (qerrName[m]) vQerr[m] = w.def_var(qerrName[m], nc3::NC_DOUBLE, {dSeg}, {{"units", "m3/s"}}); }
    int vNt = -1, vHt = -1; std::vector<int> vHist;
    if (hist) {
        const int dOne = w.def_dim("scalar", 1);
        vNt = w.def_var("nt", nc3::NC_INT, {dOne}, {{"long_name", "Number of current acculated time steps in history variable"}});
        vHt = w.def_var("history_time", nc3::NC_DOUBLE, {dTb}, {{"units", "s"}, {"long_name", "history time"}});
        int dHru = -1;
        for (const HistVar &v : hist->vars) {
            if (v.hru && dHru < 0) dHru = w.def_dim("hru", v.n);
            vHist.push_back(w.def_var(v.name, nc3::NC_DOUBLE, {v.hru ? dHru : dSeg}));
        }
    }
    w.end_def();
    w.put_int(vId, segId.data());
    if (hist) {
        w.put_int(vNt, &hist->nt); w.put_double(vHt, hist->tb);
        for (size_t i = 0; i < hist->vars.size(); ++i) w.put_double(vHist[i], hist->vars[i].data->data() + hist->vars[i].offset);
    }
    const double tb[2] = {T0 - o.dt, T0}; w.put_double(vTb, tb); (void)steps;      // T0 = TSEC(1) of the NEXT step: the file holds the last step's bounds
    auto transposed = [&](const std::vector<double> &a, int ncol) { std::vector<double> t(a.size()); for (size_t r = 0; r < N; ++r) for (int k = 0; k < ncol; ++k) t[(size_t)k * N + r] = a[r * ncol + k]; return t; };
    std::vector<double> a((size_t)N * std::max(std::max(nb, mx), W)), b(N);
    a.resize((size_t)N * 2); check(mr_get_state(h, MR_ST_BASIN_QR, a.data(), (long)a.size() * 8, msg), msg);
    for (size_t r = 0; r < N; ++r) b[r] = a[2 * r + 1];
    w.put_double(vBq, b.data());
    a.resize((size_t)N * nb); check(mr_get_state(h, MR_ST_BASIN_QFUTURE, a.data(), (long)a.size() * 8, msg), msg); w.put_double(vQf, transposed(a, nb).data());
    std::vector<double> lv((size_t)o.n_routes * N); check(mr_get_state(h, MR_ST_LAKE_VOL, lv.data(), (long)lv.size() * 8, msg), msg);
    if (irf) {
        std::vector<int> nt(N); std::vector<double> uh((size_t)N * mx);
        check(mr_get_reach_uh(h, nt.data(), uh.data(), msg), msg);
        w.put_int(vNqf, nt.data());
        a.resize((size_t)N * mx); check(mr_get_state(h, MR_ST_IRF_QFUTURE, a.data(), (long)a.size() * 8, msg), msg);
        for (size_t r = 0; r < N; ++r) for (int k = nt[r]; k < mx; ++k) a[r * mx + k] = -9999.0;          // realMissing beyond numQF (:1009)
        w.put_double(vIq, transposed(a, mx).data());
        for (int q = 0; q < o.n_routes; ++q) if (o.route_methods[q] == MR_IMPULSE_RESPONSE_FUNC) w.put_double(vIv, &lv[(size_t)q * N]);
    }
    if (kwt) {
        std::vector<int> nw(N), rt((size_t)N * W), rtT((size_t)N * W);
        check(mr_get_state(h, MR_ST_KWT_NWAVE, nw.data(), (long)N * 4, msg), msg); w.put_int(vNw, nw.data());
        check(mr_get_state(h, MR_ST_KWT_ROUTED, rt.data(), (long)rt.size() * 4, msg), msg);
        for (size_t r = 0; r < N; ++r) for (int k = 0; k < W; ++k) rtT[(size_t)k * N + r] = k < nw[r] ? rt[r * W + k] : -9999;
        w.put_int(vRt, rtT.data());
        a.resize((size_t)N * W);
        check(mr_get_state(h, MR_ST_KWT_TENTRY, a.data(), (long)a.size() * 8, msg), msg); w.put_double(vTe, transposed(a, W).data());
        check(mr_get_state(h, MR_ST_KWT_TEXIT, a.data(), (long)a.size() * 8, msg), msg); w.put_double(vTx, transposed(a, W).data());
        check(mr_get_state(h, MR_ST_KWT_QWAVE, a.data(), (long)a.size() * 8, msg), msg); w.put_double(vQw, transposed(a, W).data());
        std::fill(a.begin(), a.end(), -9999.0); w.put_double(vQm, a.data());                                  // QM is a dead field (always realMissing)
        for (int q = 0; q < o.n_routes; ++q) if (o.route_methods[q] == MR_KINEMATIC_WAVE_TRACKING) w.put_double(vKv, &lv[(size_t)q * N]);
    }
    for (int q = 0; q < o.n_routes; ++q) {                                 // Euler schemes: q_sub_<m> [mol, seg], volume_<m>
        const int m = o.route_methods[q];
        if (m < MR_KINEMATIC_WAVE) continue;
        const int nm = molN[m - 3];
        a.resize((size_t)N * nm); check(mr_get_state(h, MR_ST_MOLECULE_KW + (m - 3), a.data(), (long)a.size() * 8, msg), msg);
        w.put_double(vMol[m - 3], transposed(a, nm).data());
        w.put_double(vMolVol[m - 3], &lv[(size_t)q * N]);
    }
    if (da) {
        std::vector<double> qe((size_t)o.n_routes * N); check(mr_get_state(h, MR_ST_QERROR, qe.data(), (long)qe.size() * 8, msg), msg);
        for (int q = 0; q < o.n_routes; ++q) if (vQerr[o.route_methods[q]] >= 0) w.put_double(vQerr[o.route_methods[q]], &qe[(size_t)q * N]);
    }
    w.close();
}

// returns TSEC(1) of the next step
double read_restart(mr_handle h, const std::string &path, const mr_options &o, const std::vector<int> &segId, bool da = false, HistState *hist = nullptr) {
    char msg[MR_STRLEN];
    const size_t N = segId.size();
    nc3::Reader r(path);
    std::vector<int> id; r.read_int(r.var("reachID"), id);
    if (id != segId) die(20, "read_state_nc/reach ids of the restart file differ from the river network");
    const int nb = (int)mr_get_info(h, MR_INFO_NTDH_BAS), mx = (int)mr_get_info(h, MR_INFO_MAXTDH), W = MR_KW_SLOTS;
    if ((int)r.dim_len("tdh") != nb) die(20, "read_state_nc/tdh of the restart file differs from the hillslope UH length");
    std::vector<double> tb; r.read_all(r.var("time_bound"), tb);
    // the time bound is that of the step before the restart, the run continues one step later (init_model_data.f90:607-608)
    const long steps = std::lround(tb[1] / o.dt);
    check(mr_set_steps_done(h, steps, msg), msg);
    auto rowmajor = [&](const std::vector<double> &t, int ncol) { std::vector<double> a(t.size()); for (size_t q = 0; q < N; ++q) for (int k = 0; k < ncol; ++k) a[q * ncol + k] = t[(size_t)k * N + q]; return a; };
    std::vector<double> t, a;
    r.read_all(r.var("basin_q"), t); a.assign(2 * N, 0.0); for (size_t q = 0; q < N; ++q) a[2 * q + 1] = t[q];
    check(mr_set_state(h, MR_ST_BASIN_QR, a.data(), (long)a.size() * 8, msg), msg);
    r.read_all(r.var("qfuture"), t); a = rowmajor(t, nb); check(mr_set_state(h, MR_ST_BASIN_QFUTURE, a.data(), (long)a.size() * 8, msg), msg);
    std::vector<double> lv((size_t)o.n_routes * N, 0.0);
    for (int q = 0; q < o.n_routes; ++q) {
        if (o.route_methods[q] == MR_IMPULSE_RESPONSE_FUNC) {
            if ((int)r.dim_len("tdh_irf") != mx) die(20, "read_state_nc/tdh_irf of the restart file differs from the reach UH length");
            r.read_all(r.var("irf_qfuture"), t); a = rowmajor(t, mx);
            for (auto &v : a) if (v == -9999.0) v = 0.0;
            check(mr_set_state(h, MR_ST_IRF_QFUTURE, a.data(), (long)a.size() * 8, msg), msg);
            r.read_all(r.var("volume_irf"), t); std::copy(t.begin(), t.end(), lv.begin() + (size_t)q * N);
            check(mr_set_state(h, MR_ST_IRF_VOL, t.data(), (long)N * 8, msg), msg);
        } else if (o.route_methods[q] == MR_KINEMATIC_WAVE_TRACKING) {
            if ((int)r.dim_len("wave") != W) die(20, "read_state_nc/wave dimension of the restart file is not MR_KW_SLOTS");
            std::vector<int> nw, rt; r.read_int(r.var("numWaves"), nw); r.read_int(r.var("routed"), rt);
            std::vector<int> rr((size_t)N * W);
            for (size_t s = 0; s < N; ++s) for (int k = 0; k < W; ++k) rr[s * W + k] = (k < nw[s] && rt[(size_t)k * N + s] == 1) ? 1 : 0;
            check(mr_set_state(h, MR_ST_KWT_NWAVE, nw.data(), (long)N * 4, msg), msg);
            check(mr_set_state(h, MR_ST_KWT_ROUTED, rr.data(), (long)rr.size() * 4, msg), msg);
            r.read_all(r.var("tentry"), t); a = rowmajor(t, W); check(mr_set_state(h, MR_ST_KWT_TENTRY, a.data(), (long)a.size() * 8, msg), msg);
            r.read_all(r.var("texit"), t); a = rowmajor(t, W); check(mr_set_state(h, MR_ST_KWT_TEXIT, a.data(), (long)a.size() * 8, msg), msg);
            r.read_all(r.var("qwave"), t); a = rowmajor(t, W); check(mr_set_state(h, MR_ST_KWT_QWAVE, a.data(), (long)a.size() * 8, msg), msg);
            r.read_all(r.var("volume_kwt"), t); std::copy(t.begin(), t.end(), lv.begin() + (size_t)q * N);
        } else if (o.route_methods[q] >= MR_KINEMATIC_WAVE) {
            static const char *eul[3] = {"kw", "mc", "dw"}; static const int molN[3] = {20, 2, 20};
            const int m = o.route_methods[q]; const std::string e = eul[m - 3];
            if ((int)r.dim_len("mol_" + e) != molN[m - 3]) die(20, "read_state_nc/molecule dimension of the restart file differs");
            r.read_all(r.var("q_sub_" + e), t); a = rowmajor(t, molN[m - 3]);
            check(mr_set_state(h, MR_ST_MOLECULE_KW + (m - 3), a.data(), (long)a.size() * 8, msg), msg);
            r.read_all(r.var("volume_" + e), t); std::copy(t.begin(), t.end(), lv.begin() + (size_t)q * N);
        }
    }
    check(mr_set_state(h, MR_ST_LAKE_VOL, lv.data(), (long)lv.size() * 8, msg), msg);
    if (da) {                                   // qerror_<method> if the file has it, else 0 (read_restart.f90:358-366); Qobs / Qelapsed start from 0
        static const char *qerrName[6] = {nullptr, "qerror_irf", nullptr, "qerror_kw", "qerror_mc", "qerror_dw"};
        std::vector<double> qe((size_t)o.n_routes * N, 0.0);
        for (int q = 0; q < o.n_routes; ++q) {
            const char *nm = qerrName[o.route_methods[q]];
            if (nm && r.find(nm)) { r.read_all(r.var(nm), t); std::copy(t.begin(), t.end(), qe.begin() + (size_t)q * N); }
        }
        check(mr_set_state(h, MR_ST_QERROR, qe.data(), (long)qe.size() * 8, msg), msg);
    }
    if (hist && r.find("nt")) {                     // the open output period (absent in files of runs without aggregation)
        std::vector<int> nt; r.read_int(r.var("nt"), nt);
        hist->nt = nt.empty() ? 0 : nt[0];
        if (hist->nt > 0) {
            r.read_all(r.var("history_time"), t); hist->tb[0] = t[0]; hist->tb[1] = t[1];
            for (HistVar &v : hist->vars) {
                if (!r.find(v.name)) die(20, "read_state_nc/the restart file was written inside an output period but does not hold " + v.name);
                r.read_all(r.var(v.name), t);
                if (t.size() != v.n) die(20, "read_state_nc/size of " + v.name + " in the restart file");
                std::copy(t.begin(), t.end(), v.data->begin() + v.offset);
            }
        }
    }
    return tb[0] + o.dt;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: route_runoff <control file> [--batch N] [--device-ingest] [--device-history] [--dry-run [--dump-forcing FILE] [--dump-remap FILE]]\n"); return 2; }
    const std::string cfile = argv[1];
    int batch = 64; bool dry = false, deviceIngest = false, deviceHistory = false; std::string dumpForcing, dumpRemap;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--batch") && i + 1 < argc) batch = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--dry-run")) dry = true;
        else if (!std::strcmp(argv[i], "--device-ingest")) deviceIngest = true;       // forcing records -> runoff rows on the device (mr_ingest_records)
        else if (!std::strcmp(argv[i], "--device-history")) deviceHistory = true;     // period means of the history file on the device (mr_history_means)
        else if (!std::strcmp(argv[i], "--dump-forcing") && i + 1 < argc) dumpForcing = argv[++i];
        else if (!std::strcmp(argv[i], "--dump-remap") && i + 1 < argc) dumpRemap = argv[++i];
        else die(2, std::string("unknown argument ") + argv[i]);
    }
    if (batch < 1) die(2, "--batch must be >= 1");
    try {
        // ---- init_model: control file + parameter namelist
        const Control c = read_control(cfile);
        // options of the reference this host accepts but does not act on: say so when they ask for something non-default
        if (c.flag("tracer", false)) std::fprintf(stderr, "route_runoff: <tracer> T is ignored (solute transport is not built)\n");
        if (c.flag("outputAtGage", false) && c.str("gageMetaFile", "").empty()) std::fprintf(stderr, "route_runoff: <outputAtGage> T without <gageMetaFile> is ignored (history holds every reach)\n");
        if (c.num("seg_outlet", -9999.0) != -9999.0) std::fprintf(stderr, "route_runoff: <seg_outlet> is ignored (the whole network of <fname_ntopOld> is routed)\n");
        const std::string ancil = c.need("ancil_dir"), indir = c.need("input_dir"), outdir = c.need("output_dir");
        mr_options o{};
        o.dt = c.num("dt_qsim", -1.0);
        if (!(o.dt > 0.0)) die(20, "read_control/<dt_qsim> must be given");
        if (std::fmod(86400.0, o.dt) != 0.0 && std::fmod(o.dt, 86400.0) != 0.0) die(20, "read_control/<dt_qsim> must be a divisor or a multiple of 86400 s");
        const std::string ropt = c.need("route_opt");
        o.n_routes = 0;
        for (char ch : ropt) {                                   // read_control.f90:583-597
            if (ch < '0' || ch > '5') die(81, "read_control/route_opt must be a string of digits 0-5");
            if (o.n_routes >= 8) die(81, "read_control/too many routing methods");
            o.route_methods[o.n_routes++] = ch - '0';
        }
        o.doesBasinRoute = (int)c.num("doesBasinRoute", 1);
        o.hw_drain_point = (int)c.num("hw_drain_point", 2);
        o.min_length_route = c.num("min_length_route", 0.0);
        o.is_lake_sim = c.flag("is_lake_sim", false); o.lakeRegulate = c.flag("lakeRegulate", true); o.LakeInputOption = (int)c.num("LakeInputOption", 0);
        o.runoffMin = c.num("runoffMin", 0.0);
        const bool isRemap = c.flag("is_remap", false);
        const bool fluxWm = c.flag("is_flux_wm", false), volWm = c.flag("is_vol_wm", false) && o.is_lake_sim;    // main_route.f90:110-123
        {                                                        // units_qsim -> time_conv, length_conv (read_control.f90:443-474)
            const std::string u = c.need("units_qsim");
            const size_t sl = u.find('/');
            if (sl == std::string::npos) die(20, "read_control/expect the character \"/\" exists in the units string");
            const std::string cl = trim(u.substr(0, sl)), ct = trim(u.substr(sl + 1));
            if (cl == "m") o.length_conv = 1.0; else if (cl == "mm") o.length_conv = 1.0 / 1000.0; else die(20, "read_control/expect the length units of runoff to be m or mm");
            if (ct == "d" || ct == "day") o.time_conv = 1.0 / 86400.0; else if (ct == "h" || ct == "hr" || ct == "hour") o.time_conv = 1.0 / 3600.0;
            else if (ct == "s" || ct == "sec" || ct == "second") o.time_conv = 1.0; else die(20, "read_control/expect the time units of runoff to be day(d), hour(h) or second(s)");
        }
        o.fshape = 2.5; o.tscale = 86400.0; o.velo = 1.5; o.diff = 5000.0; o.mann_n = 0.01; o.wscale = 0.001;   // param.nml.default
        read_param_nml(join_path(ancil, c.need("param_nml")), o);
        o.device = std::getenv("MR_DEVICE") ? std::atoi(std::getenv("MR_DEVICE")) : 0;
        o.max_batch = batch;
        o.floodplain = c.flag("floodplain", false);              // Euler schemes: finite bankfull depth (process_ntopo.f90:190-196)

        // ---- init_ntopo: river network (read_streamSeg.f90:44-276)
        nc3::Reader nt(join_path(ancil, c.need("fname_ntopOld")));
        const size_t nRch = nt.dim_len(c.str("dname_sseg", "seg")), nHRU = nt.dim_len(c.str("dname_nhru", "hru"));
        std::vector<int> segId, downSegId, hruId, hruSegId, islake, lakeType;
        std::vector<double> area, length, slope, width, man_n, d03[4];
        nt.read_int(nt.var(c.str("varname_segId", "segId")), segId);
        nt.read_int(nt.var(c.str("varname_downSegId", "downSegId")), downSegId);
        nt.read_int(nt.var(c.str("varname_HRUid", "HRUid")), hruId);
        nt.read_int(nt.var(c.str("varname_hruSegId", "hruSegId")), hruSegId);
        nt.read_all(nt.var(c.str("varname_area", "area")), area);
        nt.read_all(nt.var(c.str("varname_length", "length")), length);
        nt.read_all(nt.var(c.str("varname_slope", "slope")), slope);
        const bool geomFromFile = (int)c.num("hydGeometryOption", 1) == 0;        // 0 = read width / man_n from the file
        if (geomFromFile) { nt.read_all(nt.var(c.str("varname_width", "width")), width); nt.read_all(nt.var(c.str("varname_man_n", "man_n")), man_n); }
        if (o.is_lake_sim) {
            nt.read_int(nt.var(c.str("varname_islake", "islake")), islake);
            if (const nc3::Var *v = nt.find(c.str("varname_lakeModelType", "lakeModelType"))) nt.read_int(*v, lakeType);
            const char *nm[4] = {"D03_MaxStorage", "D03_Coefficient", "D03_Power", "D03_S0"};
            for (int k = 0; k < 4; ++k) if (const nc3::Var *v = nt.find(c.str(std::string("varname_") + nm[k], nm[k]))) nt.read_all(*v, d03[k]);
        }
        if (segId.size() != nRch || hruId.size() != nHRU) die(20, "read_streamSeg/variable sizes do not match the seg/hru dimensions");

        // ---- init_inFile_pop: forcing file(s), time axis, HRU ids
        std::vector<Forcing> files;
        {
            const std::string q = join_path(indir, c.need("fname_qsim"));
            std::vector<std::string> paths;
            { FILE *f = std::fopen(q.c_str(), "rb"); if (!f) die(30, "inFile_pop/" + q + " does not exist"); unsigned char m[3] = {0, 0, 0}; const size_t got = std::fread(m, 1, 3, f); std::fclose(f);
              if (got == 3 && m[0] == 'C' && m[1] == 'D' && m[2] == 'F') paths.push_back(q);
              else { std::ifstream lst(q); std::string ln; while (std::getline(lst, ln)) { ln = trim(ln); if (!ln.empty() && ln[0] != '!') paths.push_back(join_path(indir, ln)); } } }
            if (paths.empty()) die(20, "inFile_pop/no forcing file listed in " + q);
            for (auto &p : paths) { Forcing f; f.path = p; files.push_back(f); }
        }
        const std::string vtime = c.need("vname_time"), vq = c.need("vname_qsim");
        bool noleap = false; std::vector<int> roHruId;
        // the runoff variable is [time, hru] (vector of polygons) or [time, lat, lon] (grid; needs <is_remap> T):
        // read_forcing_metadata, model_setup.f90:741-775
        bool grid = false; size_t nLat = 0, nLon = 0;
        for (auto &f : files) {
            nc3::Reader r(f.path);
            const nc3::Var &tv = r.var(vtime);
            std::string cal = c.str("ro_calendar", r.attr_text(tv, "calendar"));
            cal = lower(cal); noleap = (cal == "noleap" || cal == "365_day");
            if (!(cal.empty() || cal == "standard" || cal == "gregorian" || cal == "proleptic_gregorian" || noleap)) die(20, "inFile_pop/calendar '" + cal + "' is not supported by this host");
            double scale, epoch; parse_time_units(c.str("ro_time_units", r.attr_text(tv, "units")), noleap, scale, epoch);
            std::vector<double> tt; r.read_all(tv, tt);
            f.nTime = tt.size(); f.tsec.resize(tt.size());
            for (size_t i = 0; i < tt.size(); ++i) f.tsec[i] = epoch + tt[i] * scale;
            const size_t rank = r.var(vq).dimids.size();
            if (rank != 2 && rank != 3) die(20, "init_runoff_data/ndims of input data invalid");
            if (rank == 3) {
                const size_t la = r.dim_len(c.str("dname_ylat", "lat")), lo = r.dim_len(c.str("dname_xlon", "lon"));
                if (grid && (la != nLat || lo != nLon)) die(20, "inFile_pop/forcing files have different grids");
                grid = true; nLat = la; nLon = lo;
            } else if (roHruId.empty()) r.read_int(r.var(c.need("vname_hruid")), roHruId);
        }
        if (grid && !isRemap) die(20, "init_runoff_data/gridded runoff needs <is_remap> T and a mapping file with i_index/j_index");
        const size_t nForcing = grid ? nLat * nLon : roHruId.size();
        std::vector<double> tAll; std::vector<std::pair<int, size_t>> where;    // (file, record) of every forcing time
        for (size_t k = 0; k < files.size(); ++k) for (size_t i = 0; i < files[k].nTime; ++i) { tAll.push_back(files[k].tsec[i]); where.push_back({(int)k, i}); }
        // ---- init_time (model_setup.f90:404-566): forcing interval and period, simulation period clipped to the forcing
        double dtro = c.num("dt_ro", o.dt);                        // a single record: <dt_ro>, else <dt_qsim>
        if (tAll.size() >= 2) {
            dtro = tAll[1] - tAll[0];                              // dt_ro is taken from the file (model_setup.f90:502)
            if (!(dtro > 0.0)) die(20, "init_time/forcing times are not increasing");
            for (size_t i = 2; i < tAll.size(); ++i)               // maxTimeDiff = 1 s, public_var.f90:28
                if (std::fabs((tAll[i] - tAll[i - 1]) - dtro) > 1.0) die(20, "init_time/make sure the input netCDF files do not have time overlaps or gaps");
        }
        const std::string stampAt = lower(c.str("ro_time_stamp", "start"));
        if (stampAt != "start" && stampAt != "end" && stampAt != "middle") die(20, "read_control/Input time stamp <ro_time_stamp> must be start, end, or middle");
        const size_t nRo = tAll.size();
        if (nRo == 0) die(20, "init_time/the forcing files hold no time record");
        const double roBeg = tAll[0] - (stampAt == "end" ? dtro : stampAt == "middle" ? 0.5 * dtro : 0.0), roEnd = roBeg + (double)nRo * dtro;
        const double tStartAsked = parse_datetime(c.need("sim_start"), noleap);
        double tStart = tStartAsked, tEnd = parse_datetime(c.need("sim_end"), noleap);
        if (tEnd < tStart) die(20, "init_time/simulation end is before simulation start");
        if (tStart > roEnd) die(20, "init_time/check <sim_start> against runoff input time");
        if (tStart < roBeg) { std::fprintf(stderr, "WARNING: <sim_start> is before the first time step in input runoff; reset to runoff_start\n"); tStart = roBeg; }
        if (tEnd > roEnd) { std::fprintf(stderr, "WARNING: <sim_end> is after the last time step in input runoff; reset to runoff_end\n"); tEnd = roEnd; }
        // update_time (init_model_data.f90:284-326): the step that starts at or after <sim_end> is the last one.  Steps
        // the forcing does not cover completely are not run (the reference's map is undefined there).
        const double tolT = 1e-6;
        size_t nSteps = (size_t)std::ceil((tEnd - tStart) / o.dt - tolT) + 1;
        { const size_t nCovered = (size_t)std::floor((roEnd - tStart) / o.dt + tolT); if (nSteps > nCovered) nSteps = nCovered; }
        if (nSteps == 0) die(20, "init_time/no forcing record between <sim_start> and <sim_end>");
        // timeMap_sim_forc (get_basin_runoff.f90:256-369): forcing records under simulation step k and their weights.
        // One record (dt_qsim <= dt_ro, step inside a record): that record, no weight.
        struct TimeMap { std::vector<size_t> rec; std::vector<double> frac; };
        auto time_map_of = [&](double beg, double dtIn, size_t nRec, size_t k) {      // the same for the runoff and the water-management files
            TimeMap m;
            const double sim1 = (tStart - beg) + (double)k * o.dt, sim2 = sim1 + o.dt;
            double f0 = std::floor(sim1 / dtIn + tolT); if (f0 < 0.0) f0 = 0.0;
            size_t front = (size_t)f0;                                                   // first record whose end is after sim1
            double e = std::ceil(sim2 / dtIn - tolT) - 1.0; if (e < 0.0) e = 0.0;    // first record whose end is at or after sim2
            size_t end = (size_t)e;
            if (end > nRec - 1) end = nRec - 1;
            if (front > end) die(30, "timeMap_sim_forc/index of idxFront lower than idxEnd");
            for (size_t r = front; r <= end; ++r) {
                m.rec.push_back(r);
                if (front == end) break;
                m.frac.push_back(r == front ? ((double)(r + 1) * dtIn - sim1) / o.dt : r == end ? (sim2 - (double)r * dtIn) / o.dt : dtIn / o.dt);
            }
            return m; };
        auto time_map = [&](size_t k) { return time_map_of(roBeg, dtro, nRo, k); };
        const size_t i0 = time_map(0).rec[0];

        // sort_flux index: forcing HRU -> network HRU (process_remap.f90:271-311)
        std::vector<int> ix(nForcing, -1);
        { std::vector<std::pair<int, int>> tab(nHRU); for (size_t i = 0; i < nHRU; ++i) tab[i] = {hruId[i], (int)i}; std::sort(tab.begin(), tab.end());
          for (size_t i = 0; i < roHruId.size(); ++i) { auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(roHruId[i], -1)); if (it != tab.end() && it->first == roHruId[i]) ix[i] = it->second; } }

        std::printf("{\"case\": \"%s\", \"nRch\": %zu, \"nHRU\": %zu, \"nHRU_forcing\": %zu, \"nSteps\": %zu, \"dt\": %.1f, \"route_opt\": \"%s\", \"first_record\": %zu, "
                    "\"fshape\": %.6g, \"tscale\": %.6g, \"velo\": %.6g, \"diff\": %.6g, \"mann_n\": %.6g, \"wscale\": %.6g, \"time_conv\": %.9g, \"length_conv\": %.9g}\n",
                    c.str("case_name", "case").c_str(), nRch, nHRU, nForcing, nSteps, o.dt, ropt.c_str(), i0, o.fshape, o.tscale, o.velo, o.diff, o.mann_n, o.wscale, o.time_conv, o.length_conv);
        // <outputFrequency>: every step, a number of steps, or "daily"; fluxes are averaged over the period
        // (histVars_data.f90:154-246) and stamped with the start of the period
        int nAgg = 1;
        { const std::string of = lower(c.str("outputFrequency", "1"));
          if (of == "daily") { if (std::fmod(86400.0, o.dt) != 0.0) die(20, "route_runoff/<outputFrequency> daily needs dt_qsim to divide 86400 s"); nAgg = (int)std::lround(86400.0 / o.dt); }
          else { char *e; const long v = std::strtol(of.c_str(), &e, 10); if (e == of.c_str() || v < 1) die(20, "route_runoff/<outputFrequency> " + of + ": only an integer number of steps or 'daily' is supported by this host"); nAgg = (int)v; } }
        // <newFileFrequency>: which history file every output record goes to
        const std::string fileFreq = lower(c.str("newFileFrequency", "single"));
        if (fileFreq != "single" && fileFreq != "daily" && fileFreq != "monthly" && fileFreq != "yearly") die(20, "new_file_alarm/unable to identify the option to define new output files");
        struct HistFile { std::string path; size_t first, nrec; };          // first step held, records written
        std::vector<HistFile> plan;
        { int na = 0;
          for (size_t k = 0; k < nSteps; ++k) {
              const Civil cur = civil_from_sec(tStart + (double)k * o.dt, noleap);
              if (k == 0 || new_file_alarm(fileFreq, civil_from_sec(tStart + (double)(k - 1) * o.dt, noleap), cur))
                  plan.push_back({join_path(outdir, c.str("case_name", "case") + ".h." + hist_stamp(fileFreq, cur) + ".nc"), k, 0});
              if (++na == nAgg || k + 1 == nSteps) { ++plan.back().nrec; na = 0; }
          } }

        // <restart_write>: the steps after which the state is written (restart_alarm, write_restart_pio.f90:110-163; init_time,
        // model_setup.f90:646-686).  The file is stamped with the start of the NEXT step (restart_fname, :207-253), which is
        // what the periodic options compare with <restart_month> / <restart_day> / <restart_hour>.
        const std::string rw = lower(c.str("restart_write", "never"));
        std::vector<std::pair<size_t, std::string>> restartPlan;             // (last step of the state, file)
        {
            if (rw != "never" && rw != "last" && rw != "specified" && rw != "yearly" && rw != "monthly" && rw != "daily")
                die(20, "init_time/Accepted <restart_write> options: last, never, specified, yearly, monthly, or daily");
            const int rMon = (int)c.num("restart_month", 1), rDay = (int)c.num("restart_day", 1), rHour = (int)c.num("restart_hour", 0);
            double tSpec = 0.0;
            if (rw == "specified") { if (!c.has("restart_date")) die(20, "init_time/<restart_date> must be provided when <restart_write> option is \"specified\""); tSpec = parse_datetime(c.need("restart_date"), noleap); }
            auto ndays = [&](int y, int m) { static const int ml[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
                                             return ml[m - 1] + ((m == 2 && !noleap && ((y % 4 == 0 && y % 100 != 0) || y % 400 == 0)) ? 1 : 0); };
            for (size_t k = 0; rw != "never" && k < nSteps; ++k) {
                const double tNext = tStart + (double)(k + 1) * o.dt;
                const Civil cv = civil_from_sec(tNext, noleap);
                const bool atHour = cv.sod == rHour * 3600, atDay = cv.d == std::min(rDay, ndays(cv.y, cv.mo));
                const bool ring = rw == "last" ? k + 1 == nSteps : rw == "specified" ? std::fabs(tNext - tSpec) < 1e-3
                                : rw == "daily" ? atHour : rw == "monthly" ? (atHour && atDay) : (atHour && atDay && cv.mo == rMon);
                if (ring) { char rs[64]; std::snprintf(rs, sizeof rs, "%04d-%02d-%02d-%05d", cv.y, cv.mo, cv.d, cv.sod);
                            restartPlan.push_back({k, join_path(c.str("restart_dir", outdir), c.str("case_name", "case") + ".r." + rs + ".nc")}); }
            }
        }

        // ---- runoff remapping (<is_remap> T): mapping netCDF -> index form for the device-side remap (read_remap.f90:20-170,
        // process_remap.f90:59-262).  A gridded forcing is the same weighted sum over a flattened [lat][lon] record.
        std::vector<int> mapHruIx, mapNumQ, mapQIx; std::vector<double> mapWgt;
        if (isRemap) {
            nc3::Reader rm(join_path(ancil, c.need("fname_remap")));
            std::vector<int> mapId, numQ, qId; std::vector<double> wgt;
            rm.read_int(rm.var(c.need("vname_hruid_in_remap")), mapId);
            rm.read_int(rm.var(c.need("vname_num_qhru")), numQ);
            rm.read_all(rm.var(c.need("vname_weight")), wgt);
            std::vector<int> qIxGrid;
            if (grid) {                                                     // remap_2D_runoff (process_remap.f90:59-162): sim2d(ii, jj), ii along lon, jj along lat, 1-based;
                std::vector<int> ii, jj;                                    // a cell outside the grid is skipped like an unknown polygon id
                rm.read_int(rm.var(c.str("vname_i_index", "i_index")), ii);
                rm.read_int(rm.var(c.str("vname_j_index", "j_index")), jj);
                if (ii.size() != wgt.size() || jj.size() != wgt.size()) die(20, "read_remap/mapping variables have inconsistent sizes");
                qIxGrid.resize(ii.size());
                for (size_t i = 0; i < ii.size(); ++i)
                    qIxGrid[i] = (ii[i] < 1 || ii[i] > (int)nLon || jj[i] < 1 || jj[i] > (int)nLat) ? -1 : (jj[i] - 1) * (int)nLon + (ii[i] - 1);
            } else rm.read_int(rm.var(c.need("vname_qhruid")), qId);
            if (mapId.size() != numQ.size() || (!grid && qId.size() != wgt.size())) die(20, "read_remap/mapping variables have inconsistent sizes");
            auto lookup = [](const std::vector<int> &keys, const std::vector<int> &ids) {
                std::vector<std::pair<int, int>> tab(ids.size()); for (size_t i = 0; i < ids.size(); ++i) tab[i] = {ids[i], (int)i}; std::sort(tab.begin(), tab.end());
                std::vector<int> out(keys.size(), -1);
                for (size_t i = 0; i < keys.size(); ++i) { auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(keys[i], -1)); if (it != tab.end() && it->first == keys[i]) out[i] = it->second; }
                return out; };
            mapHruIx = lookup(mapId, hruId); mapQIx = grid ? qIxGrid : lookup(qId, roHruId); mapNumQ = numQ; mapWgt = wgt;
            if (!dumpRemap.empty()) {                                       // int32 n, m, hruIx[n], numQ[n], qIx[m]; float64 w[m] -- what mr_set_remap receives
                FILE *f = std::fopen(dumpRemap.c_str(), "wb"); if (!f) die(30, "route_runoff/cannot write " + dumpRemap);
                const int nm[2] = {(int)mapHruIx.size(), (int)mapQIx.size()};
                std::fwrite(nm, sizeof(int), 2, f); std::fwrite(mapHruIx.data(), sizeof(int), mapHruIx.size(), f); std::fwrite(mapNumQ.data(), sizeof(int), mapNumQ.size(), f);
                std::fwrite(mapQIx.data(), sizeof(int), mapQIx.size(), f); std::fwrite(mapWgt.data(), sizeof(double), mapWgt.size(), f); std::fclose(f);
            }
        }


        // get_hru_runoff for simulation step k (get_basin_runoff.f90:19-120): the step's forcing record(s) -> one row of
        // the library's input (network HRU order, or forcing-polygon order when the device remaps)
        const size_t inCols = isRemap ? nForcing : nHRU;
        std::vector<nc3::Reader *> rd(files.size(), nullptr);
        std::vector<double> rec, wsum, wtot;
        double fillv = c.num("input_fillvalue", -9999.0);
        // One forcing variable of simulation step k -> one row.  <scale_factor_*> / <offset_value_*>: scale_forcing
        // (get_basin_runoff.f90:375-423); both ~0 = the variable is not read at all and is zero (:71-73,138-141,168-171).
        struct VarSpec { std::string name; double scale, offset; bool flip; };
        auto load_var = [&](const VarSpec &vs, size_t k, double *dst, size_t cols, bool forDeviceRemap) {
            const bool rescale = vs.scale != -9999.0 || vs.offset != -9999.0;
            const bool zero = std::fabs(vs.scale) < 2.3e-308 && (std::fabs(vs.offset) < 2.3e-308 || vs.offset == -9999.0);
            const double A = vs.scale == -9999.0 ? 1.0 : vs.scale, B = vs.offset == -9999.0 ? 0.0 : vs.offset;
            if (zero) { std::fill(dst, dst + cols, 0.0); return; }
            const TimeMap tm = time_map(k);
            for (size_t j = 0; j < tm.rec.size(); ++j) {
                const auto wr = where[tm.rec[j]];
                if (!rd[wr.first]) rd[wr.first] = new nc3::Reader(files[wr.first].path);
                nc3::Reader &R = *rd[wr.first];
                const nc3::Var &qv = R.var(vs.name);
                double fv; if (R.attr_value(qv, "_FillValue", fv)) fillv = fv;
                R.read(qv, rec, wr.second, 1);
                if (rec.size() != nForcing) die(20, "read_runoff/forcing variable " + vs.name + " is not dimensioned like the runoff");
                if (tm.frac.empty()) break;
                // several records under one step: time-weighted mean over the records that hold a value
                // (read_1D_forcing, read_runoff.f90:298-325)
                if (j == 0) { wsum.assign(rec.size(), 0.0); wtot.assign(rec.size(), 0.0); }
                for (size_t i = 0; i < rec.size(); ++i) if (rec[i] != fillv) { wsum[i] += rec[i] * tm.frac[j]; wtot[i] += tm.frac[j]; }
            }
            if (!tm.frac.empty())
                for (size_t i = 0; i < rec.size(); ++i) rec[i] = wtot[i] == 0.0 ? fillv : (wtot[i] < 1.0 ? wsum[i] / wtot[i] : wsum[i]);
            if (vs.flip) for (size_t i = 0; i < rec.size(); ++i) if (rec[i] != fillv && rec[i] != -9999.0) rec[i] = -1.0 * rec[i] + 0.0;   // <is_Ep_upward_negative>
            if (rescale) for (size_t i = 0; i < rec.size(); ++i) if (rec[i] != fillv && rec[i] != -9999.0) rec[i] = A * rec[i] + B;
            if (forDeviceRemap) { for (size_t i = 0; i < rec.size(); ++i) dst[i] = rec[i] == fillv ? -9999.0 : rec[i]; return; }  // remapped on the device; realMissing (< 0) is skipped there
            std::fill(dst, dst + nHRU, 0.0);                                 // HRUs without forcing: realMissing -> 0 (sort_flux)
            for (size_t i = 0; i < rec.size(); ++i) if (ix[i] >= 0) { double v = rec[i]; if (v == fillv || v < 0.0) v = 0.0; dst[ix[i]] = v; }
        };
        const VarSpec vsRunoff{vq, c.num("scale_factor_runoff", -9999.0), c.num("offset_value_runoff", -9999.0), false};
        auto load_step = [&](size_t k, double *dst) { load_var(vsRunoff, k, dst, inCols, isRemap); };
        // lake evaporation / precipitation (get_basin_runoff.f90:136-197): read with <is_lake_sim> T and <LakeInputOption> 0 or 2
        const bool lakeForcing = o.is_lake_sim && (o.LakeInputOption == 0 || o.LakeInputOption == 2);
        if (lakeForcing && isRemap) die(20, "route_runoff/lake evaporation and precipitation with <is_remap> T are not supported by this host");
        const VarSpec vsEvapo{c.str("vname_evapo", "evapo"), c.num("scale_factor_Ep", -9999.0), c.num("offset_value_Ep", -9999.0), c.flag("is_Ep_upward_negative", false)};
        const VarSpec vsPrecip{c.str("vname_precip", "precip"), c.num("scale_factor_prec", -9999.0), c.num("offset_value_prec", -9999.0), false};

        // water management (<is_flux_wm>, <is_vol_wm>): one netCDF with [time, seg] variables (get_basin_runoff.f90:199-250);
        // sort_flux by reach id: reaches the file does not hold get realMissing (= no flux) / 0 (target volume)
        std::unique_ptr<nc3::Reader> wmFile;
        std::vector<int> wmIx; size_t nWm = 0; double wmBeg = 0.0, dtWm = o.dt; std::vector<double> wmRec, wmSum, wmTot;
        if (fluxWm || volWm) {
            wmFile.reset(new nc3::Reader(join_path(indir, c.need("fname_wm"))));
            const nc3::Var &tv = wmFile->var(c.need("vname_time_wm"));
            double scale, epoch; parse_time_units(wmFile->attr_text(tv, "units"), noleap, scale, epoch);
            std::vector<double> tt; wmFile->read_all(tv, tt);
            if (tt.empty()) die(20, "init_time/the water-management file holds no time record");
            nWm = tt.size(); wmBeg = epoch + tt[0] * scale; dtWm = nWm >= 2 ? (tt[1] - tt[0]) * scale : c.num("dt_wm", o.dt);
            if (tStart < wmBeg - 1e-3 || tStart + (double)nSteps * o.dt > wmBeg + (double)nWm * dtWm + 1e-3) die(20, "init_time/the water-management file does not cover the simulation period");
            std::vector<int> wmSeg; wmFile->read_int(wmFile->var(c.need("vname_segid_wm")), wmSeg);
            std::vector<std::pair<int, int>> tab(nRch); for (size_t i = 0; i < nRch; ++i) tab[i] = {segId[i], (int)i}; std::sort(tab.begin(), tab.end());
            wmIx.assign(wmSeg.size(), -1);
            for (size_t i = 0; i < wmSeg.size(); ++i) { auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(wmSeg[i], -1)); if (it != tab.end() && it->first == wmSeg[i]) wmIx[i] = it->second; }
        }
        auto load_wm = [&](const std::string &vname, bool removeNegatives, size_t k, double *dst) {
            const TimeMap tm = time_map_of(wmBeg, dtWm, nWm, k);
            const nc3::Var &qv = wmFile->var(vname);
            double fv = -9999.0; wmFile->attr_value(qv, "_FillValue", fv);
            for (size_t j = 0; j < tm.rec.size(); ++j) {
                wmFile->read(qv, wmRec, tm.rec[j], 1);
                if (wmRec.size() != wmIx.size()) die(20, "read_runoff/water-management variable " + vname + " is not dimensioned [time, seg]");
                if (tm.frac.empty()) break;
                if (j == 0) { wmSum.assign(wmRec.size(), 0.0); wmTot.assign(wmRec.size(), 0.0); }
                for (size_t i = 0; i < wmRec.size(); ++i) if (wmRec[i] != fv) { wmSum[i] += wmRec[i] * tm.frac[j]; wmTot[i] += tm.frac[j]; }
            }
            if (!tm.frac.empty())
                for (size_t i = 0; i < wmRec.size(); ++i) wmRec[i] = wmTot[i] == 0.0 ? fv : (wmTot[i] < 1.0 ? wmSum[i] / wmTot[i] : wmSum[i]);
            std::fill(dst, dst + nRch, removeNegatives ? 0.0 : -9999.0);
            for (size_t i = 0; i < wmRec.size(); ++i) if (wmIx[i] >= 0) { double v = wmRec[i] == fv ? -9999.0 : wmRec[i]; if (removeNegatives && v < 0.0) v = 0.0; dst[wmIx[i]] = v; }
        };

        // data assimilation (<qmodOption> 1, init_model_data.f90:841-871): gauge metadata csv <gageMetaFile> (header row; columns
        // gage_id, reach_id: gageMeta_data.f90:58-68) and the gauge netCDF <fname_gageObs> ([time, site] flow, site names as a
        // character variable: obs_data.f90:272-516), both in <ancil_dir>; a step sees the record whose time equals the start of
        // the step (gage_obs_data%time_ix(simDatetime(1)), main_route.f90:128); without the file the option is switched off
        // gauge metadata csv: (gage_id, reach_id) rows in file order (gageMeta_data.f90:58-68)
        auto read_gage_meta = [&]() {
            std::vector<std::pair<std::string, int>> rows;
            std::ifstream csv(join_path(ancil, c.need("gageMetaFile")));
            if (!csv) die(20, "read_gage_meta/cannot open " + join_path(ancil, c.need("gageMetaFile")));
            std::string line; int cg = -1, cr = -1;
            auto cells = [](const std::string &ln) { std::vector<std::string> v; std::stringstream ss(ln); std::string x; while (std::getline(ss, x, ',')) v.push_back(trim(x)); return v; };
            if (std::getline(csv, line)) { const auto hd = cells(line); for (size_t i = 0; i < hd.size(); ++i) { if (hd[i] == "gage_id") cg = (int)i; if (hd[i] == "reach_id") cr = (int)i; } }
            if (cg < 0 || cr < 0) die(20, "read_gage_meta/the csv needs the columns gage_id and reach_id");
            while (std::getline(csv, line)) { const auto v = cells(line); if ((int)v.size() > std::max(cg, cr) && !v[cg].empty()) rows.push_back({v[cg], std::atoi(v[cr].c_str())}); }
            return rows; };
        // <outputAtGage> T: the history files hold the gauged reaches only, in the order of the csv (reach_subset,
        // process_gage_meta.f90:39-86; historyFile.f90:186-199)
        std::vector<int> gageIdx;
        const bool atGage = c.flag("outputAtGage", false) && !c.str("gageMetaFile", "").empty();
        if (atGage) {
            std::vector<std::pair<int, int>> tab(nRch); for (size_t i = 0; i < nRch; ++i) tab[i] = {segId[i], (int)i}; std::sort(tab.begin(), tab.end());
            for (const auto &row : read_gage_meta()) { auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(row.second, -1)); if (it != tab.end() && it->first == row.second) gageIdx.push_back(it->second); }
            if (gageIdx.empty()) die(20, "reach_subset/<outputAtGage> T but no gauge of <gageMetaFile> lies on the river network");
        }
        int qmodOption = (int)c.num("qmodOption", 0);
        std::unique_ptr<nc3::Reader> obsFile;
        std::vector<int> obsIx; std::vector<double> obsTime, obsRec; double obsFill = -9999.0; const nc3::Var *obsVar = nullptr;
        if (qmodOption != 0 && qmodOption != 1) die(1, "init_qmod/Error: qmodOption invalid");
        if (qmodOption == 1) {
            const std::string obsPath = join_path(ancil, c.str("fname_gageObs", ""));
            FILE *probe = c.str("fname_gageObs", "").empty() ? nullptr : std::fopen(obsPath.c_str(), "rb");
            if (!probe) qmodOption = 0;
            else {
                std::fclose(probe);
                std::map<std::string, int> reachOfGage;                      // gage_id -> reach_id
                for (const auto &row : read_gage_meta()) reachOfGage[row.first] = row.second;
                obsFile.reset(new nc3::Reader(obsPath));
                const nc3::Var &tv = obsFile->var(c.need("vname_gageTime"));
                double scale, epoch; parse_time_units(obsFile->attr_text(tv, "units"), noleap, scale, epoch);
                obsFile->read_all(tv, obsTime);
                for (auto &t : obsTime) t = epoch + t * scale;
                std::vector<double> chars; const nc3::Var &sv = obsFile->var(c.need("vname_gageSite"));
                obsFile->read_all(sv, chars);
                const int dSite = obsFile->dim_index(c.need("dname_gageSite"));
                if (dSite < 0) die(20, "gageObs/dimension " + c.need("dname_gageSite") + " not found");
                const size_t nSite = obsFile->dims[dSite].len, len = nSite ? chars.size() / nSite : 0;
                std::vector<std::pair<int, int>> tab(nRch); for (size_t i = 0; i < nRch; ++i) tab[i] = {segId[i], (int)i}; std::sort(tab.begin(), tab.end());
                obsIx.assign(nSite, -1);                                   // comp_link: site -> gage_id -> reach_id -> reach index
                for (size_t i = 0; i < nSite; ++i) {
                    std::string nm; for (size_t k = 0; k < len; ++k) { const char ch = (char)chars[i * len + k]; if (ch == '\0') break; nm.push_back(ch); }
                    auto g = reachOfGage.find(trim(nm)); if (g == reachOfGage.end()) continue;
                    auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(g->second, -1));
                    if (it != tab.end() && it->first == g->second) obsIx[i] = it->second;
                }
                obsVar = &obsFile->var(c.need("vname_gageFlow"));
                obsFile->attr_value(*obsVar, "_FillValue", obsFill);
            }
        }
        // gauge row of the step that starts at absolute time t: false = no record at that time
        auto load_obs = [&](double t, double *dst) {
            for (size_t i = 0; i < nRch; ++i) dst[i] = std::nan("");
            size_t j = 0; while (j < obsTime.size() && std::fabs(obsTime[j] - t) > 0.5) ++j;     // first match, obs_data.f90:755-760
            if (j == obsTime.size()) return false;
            if (obsVar->record) obsFile->read(*obsVar, obsRec, j, 1);
            else { std::vector<double> all; obsFile->read(*obsVar, all); obsRec.assign(all.begin() + j * obsIx.size(), all.begin() + (j + 1) * obsIx.size()); }
            if (obsRec.size() != obsIx.size()) die(20, "gageObs/read_obs: the flow variable is not dimensioned [time, site]");
            for (size_t i = 0; i < obsRec.size(); ++i) if (obsIx[i] >= 0) dst[obsIx[i]] = obsRec[i] == obsFill ? std::nan("") : obsRec[i];
            return true;
        };

        if (dry) {                                                       // the time map of the first steps, for inspection
            std::printf("{\"dt_ro\": %.3f, \"ro_time_stamp\": \"%s\", \"time_map\": [", dtro, stampAt.c_str());
            for (size_t k = 0; k < std::min<size_t>(nSteps, 6); ++k) {
                const TimeMap m = time_map(k);
                std::printf("%s[", k ? ", " : "");
                for (size_t j = 0; j < m.rec.size(); ++j) std::printf("%s[%zu, %.9g]", j ? ", " : "", m.rec[j], m.frac.empty() ? 1.0 : m.frac[j]);
                std::printf("]");
            }
            std::printf("], \"history_plan\": [");
            for (size_t i = 0; i < plan.size(); ++i) std::printf("%s[\"%s\", %zu]", i ? ", " : "", plan[i].path.substr(plan[i].path.find_last_of('/') + 1).c_str(), plan[i].nrec);
            std::printf("], \"restart_plan\": [");
            for (size_t i = 0; i < restartPlan.size(); ++i) std::printf("%s[%zu, \"%s\"]", i ? ", " : "", restartPlan[i].first, restartPlan[i].second.substr(restartPlan[i].second.find_last_of('/') + 1).c_str());
            std::printf("]}\n");
            if (!dumpForcing.empty()) {                                   // raw float64 [nSteps][columns]: what the time loop would feed the library
                FILE *f = std::fopen(dumpForcing.c_str(), "wb"); if (!f) die(30, "route_runoff/cannot write " + dumpForcing);
                std::vector<double> row(inCols);
                for (size_t k = 0; k < nSteps; ++k) { load_step(k, row.data()); if (std::fwrite(row.data(), sizeof(double), inCols, f) != inCols) die(30, "route_runoff/short write to " + dumpForcing); }
                std::fclose(f);
            }
            return 0;
        }

        // ---- device side
        char msg[MR_STRLEN];
        mr_handle h = nullptr;
        int ierr = mr_create(&o, &h, msg); if (ierr) die(ierr, msg);
        if (o.is_lake_sim) {                                  // HYPE / Hanasaki reservoirs: HYP_* and H06_* of the river-network file
            std::vector<std::string> hypNames = {"HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr", "HYP_Qrate_prim",
                                                 "HYP_Qrate_amp", "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode",
                                                 "H06_Smax", "H06_alpha", "H06_envfact", "H06_S_ini", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator",
                                                 "H06_c_compare", "H06_frac_Sdead", "H06_E_rel_ini", "H06_purpose", "H06_I_mem_F", "H06_D_mem_F",
                                                 "H06_I_mem_L", "H06_D_mem_L", "LakeTargVol"};
            for (const char *mo : {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"}) {
                hypNames.push_back(std::string("H06_I_") + mo); hypNames.push_back(std::string("H06_D_") + mo); }
            for (const std::string &nmS : hypNames) {
                const char *nm = nmS.c_str();
                if (const nc3::Var *v = nt.find(c.str(std::string("varname_") + nm, nm))) {
                    std::vector<double> vals; nt.read_all(*v, vals);
                    if (vals.size() != nRch) die(20, std::string("read_streamSeg/") + nm + " is not dimensioned by segment");
                    ierr = mr_set_lake_param(h, nm, (int)nRch, vals.data(), msg); if (ierr) die(ierr, msg);
                }
            }
        }
        ierr = mr_set_network(h, (int)nRch, (int)nHRU, segId.data(), downSegId.data(), hruSegId.data(), area.data(), length.data(), slope.data(),
                              geomFromFile ? width.data() : nullptr, geomFromFile ? man_n.data() : nullptr, islake.empty() ? nullptr : islake.data(),
                              lakeType.empty() ? nullptr : lakeType.data(), d03[0].empty() ? nullptr : d03[0].data(), d03[1].empty() ? nullptr : d03[1].data(),
                              d03[2].empty() ? nullptr : d03[2].data(), d03[3].empty() ? nullptr : d03[3].data(), msg);
        if (ierr) die(ierr, msg);
        if (qmodOption == 1) { ierr = mr_set_da(h, 1, (int)c.num("qBlendPeriod", 10), (int)c.num("QerrTrend", 1), msg); if (ierr) die(ierr, msg); }
        // --device-ingest: the runoff records of a batch travel as they are in the file; the time-weighted mean over the records
        // under a step, <scale_factor_runoff> / <offset_value_runoff> and sort_flux run on the device (mr_ingest_records)
        std::vector<double> ingRec, ingFrac; std::vector<int> ingPtr, ingIdx;
        if (deviceIngest) {
            if (isRemap) die(20, "route_runoff/--device-ingest maps forcing HRUs one to one; not with <is_remap> T");
            const bool zero = std::fabs(vsRunoff.scale) < 2.3e-308 && (std::fabs(vsRunoff.offset) < 2.3e-308 || vsRunoff.offset == -9999.0);
            if (zero) deviceIngest = false;                              // the runoff is switched off: rows of zeros, nothing to ingest
        }
        if (deviceIngest) {
            std::vector<int> colOfHru(nHRU, -1);
            for (size_t i = 0; i < nForcing; ++i) if (ix[i] >= 0) colOfHru[ix[i]] = (int)i;              // sort_flux inverted (the last one wins, as there)
            const auto w0 = where[time_map(0).rec[0]];
            if (!rd[w0.first]) rd[w0.first] = new nc3::Reader(files[w0.first].path);
            double fv; if (rd[w0.first]->attr_value(rd[w0.first]->var(vsRunoff.name), "_FillValue", fv)) fillv = fv;
            ierr = mr_set_ingest(h, (int)nForcing, colOfHru.data(), vsRunoff.scale, vsRunoff.offset, fillv, msg); if (ierr) die(ierr, msg);
        }
        auto ingest_batch = [&](size_t s, int nb) {
            ingPtr.assign(1, 0); ingIdx.clear(); ingFrac.clear();
            size_t lo = (size_t)-1, hi = 0;
            for (int k = 0; k < nb; ++k) {
                const TimeMap tm = time_map(s + k);
                for (size_t j = 0; j < tm.rec.size(); ++j) {
                    ingIdx.push_back((int)tm.rec[j]); ingFrac.push_back(tm.frac.empty() ? 1.0 : tm.frac[j]);
                    lo = std::min(lo, tm.rec[j]); hi = std::max(hi, tm.rec[j]);
                }
                ingPtr.push_back((int)ingIdx.size());
            }
            for (auto &v : ingIdx) v -= (int)lo;
            ingRec.resize((hi - lo + 1) * nForcing);
            for (size_t r = lo; r <= hi; ++r) {
                const auto wr = where[r];
                if (!rd[wr.first]) rd[wr.first] = new nc3::Reader(files[wr.first].path);
                rd[wr.first]->read(rd[wr.first]->var(vsRunoff.name), rec, wr.second, 1);
                if (rec.size() != nForcing) die(20, "read_runoff/forcing variable " + vsRunoff.name + " is not dimensioned like the runoff");
                std::copy(rec.begin(), rec.end(), ingRec.begin() + (r - lo) * nForcing);
            }
            int e = mr_ingest_records(h, nb, (int)(hi - lo + 1), ingRec.data(), ingPtr.data(), ingIdx.data(), ingFrac.data(), msg); if (e) die(e, msg);
        };

        if (isRemap) { ierr = mr_set_remap(h, (int)nForcing, (int)mapHruIx.size(), mapHruIx.data(), mapNumQ.data(), mapQIx.data(), mapWgt.data(), msg); if (ierr) die(ierr, msg); }

        // ---- history files (write_simoutput_pio.f90: one float32 variable per active routing method, [time, seg])
        const char *vname[6] = {"sumUpstreamRunoff", "IRFroutedRunoff", "KWTroutedRunoff", "KWroutedRunoff", "MCroutedRunoff", "DWroutedRunoff"};
        const char *lname[6] = {"accumulated runoff from all upstream reaches", "routed runoff in each reach-impulse response function", "routed runoff in each reach-kinematic wave tracking",
                                "routed runoff in each reach-kinematic wave", "routed runoff in each reach-muskingum-cunge", "routed runoff in each reach-diffusive wave"};
        const bool wantDlay = c.flag("dlayRunoff", true);
        // Per-method reach volume (value at the end of the output period), mean inflow from upstream and the mean instantaneous
        // runoff (histVars_data.f90:200-246).  The library keeps these for the last step of a call only, so a batch ends where an
        // output period ends when a volume is asked for, and is one step long when an inflow or instRunoff is.  All default to F
        // here (the reference writes IRFvolume by default, popMetadat.f90:246).
        const char *volName[6] = {nullptr, "IRFvolume", "KWTvolume", "KWvolume", "MCvolume", "DWvolume"};
        const char *infName[6] = {nullptr, "IRFinflow", "KWTinflow", "KWinflow", "MCinflow", "DWinflow"};
        std::vector<char> wantVol(o.n_routes, 0), wantInf(o.n_routes, 0);
        bool anyVol = false, anyStep = c.flag("instRunoff", false);
        const bool wantInst = anyStep;
        for (int r = 0; r < o.n_routes; ++r) {
            const int m = o.route_methods[r];
            if (volName[m] && c.flag(volName[m], false)) { wantVol[r] = 1; anyVol = true; }
            if (infName[m] && c.flag(infName[m], false)) { wantInf[r] = 1; anyStep = true; }
        }
        // <basRunoff>: HRU runoff as it enters basin2reach, period mean, [time, hru] with basinID (historyFile.f90:156-165, default T as
        // in the reference); with <is_remap> T the remapped values exist on the device only and the variable is left out
        const bool wantBas = c.flag("basRunoff", true) && !isRemap;
        int vBas = -1;
        std::vector<int> vVol(o.n_routes, -1), vInf(o.n_routes, -1); int vInst = -1;
        int vTime = -1, vDlay = -1, vTb = -1;
        const double stampOffset = c.num("histTimeStamp_offset", 0.0);          // <histTimeStamp_offset> [s] from the start of the period (write_time, historyFile.f90:367)
        std::vector<int> vQ(o.n_routes, -1);
        std::unique_ptr<nc3::Writer> w;
        auto open_history = [&](const std::string &path) {
            if (w) w->close();
            w.reset(new nc3::Writer(path));
            const int dTime = w->def_dim("time", 0), dSeg = w->def_dim("seg", atGage ? gageIdx.size() : nRch);
            vTime = w->def_var("time", nc3::NC_DOUBLE, {dTime}, {{"units", "seconds since " + c.need("sim_start")}, {"calendar", noleap ? "noleap" : "standard"}});
            const int dTb = w->def_dim("tbound", 2);                          // time_bounds: end points of the output period (historyFile.f90:146-153)
            vTb = w->def_var("time_bounds", nc3::NC_DOUBLE, {dTime, dTb}, {{"units", "seconds since " + c.need("sim_start")}, {"long_name", "time interval endpoints"}});
            const int vId = w->def_var("reachID", nc3::NC_INT, {dSeg}, {{"long_name", "reach ID"}});
            int dHru = -1, vHid = -1;
            if (wantBas) { dHru = w->def_dim("hru", nHRU); vHid = w->def_var("basinID", nc3::NC_INT, {dHru}, {{"long_name", "basin ID"}});
                           vBas = w->def_var("basRunoff", nc3::NC_FLOAT, {dTime, dHru}, {{"units", c.need("units_qsim")}, {"long_name", "basin runoff"}}); }
            for (int r = 0; r < o.n_routes; ++r) if (c.flag(vname[o.route_methods[r]], true))
                vQ[r] = w->def_var(vname[o.route_methods[r]], nc3::NC_FLOAT, {dTime, dSeg}, {{"units", "m3/s"}, {"long_name", lname[o.route_methods[r]]}});
            if (wantDlay) vDlay = w->def_var("dlayRunoff", nc3::NC_FLOAT, {dTime, dSeg}, {{"units", "m3/s"}, {"long_name", "delayed runoff in each reach"}});
            if (wantInst) vInst = w->def_var("instRunoff", nc3::NC_FLOAT, {dTime, dSeg}, {{"units", "m3/s"}, {"long_name", "instantaneous runoff into stream or lake"}});
            for (int r = 0; r < o.n_routes; ++r) {
                const int m = o.route_methods[r];
                if (wantVol[r]) vVol[r] = w->def_var(volName[m], nc3::NC_FLOAT, {dTime, dSeg}, {{"units", "m3"}, {"long_name", "Volume in lake or stream"}});
                if (wantInf[r]) vInf[r] = w->def_var(infName[m], nc3::NC_FLOAT, {dTime, dSeg}, {{"units", "m3/s"}, {"long_name", "Inflow from upstream lake or streams"}});
            }
            w->global_attr("title", "mizuRoute routing (mizuroute-b200)");
            w->end_def();
            if (atGage) { std::vector<int> ids(gageIdx.size()); for (size_t i = 0; i < ids.size(); ++i) ids[i] = segId[gageIdx[i]]; w->put_int(vId, ids.data()); }
            else w->put_int(vId, segId.data());
            if (vHid >= 0) w->put_int(vHid, hruId.data());
        };

        // ---- time loop (route_runoff.f90:80-106), `batch` steps per library call
        std::vector<double> ro((size_t)batch * inCols), q((size_t)o.n_routes * batch * nRch);
        std::vector<double> evRows(lakeForcing ? (size_t)batch * nHRU : 0), prRows(evRows.size());
        std::vector<double> wmFluxRows(fluxWm ? (size_t)batch * nRch : 0), wmVolRows(volWm ? (size_t)batch * nRch : 0);
        std::vector<double> obsRows(qmodOption == 1 ? (size_t)batch * nRch : 0); std::vector<int> obsHas(batch, 0);
        std::vector<double> qd(wantDlay ? (size_t)batch * nRch : 0), acc((size_t)(o.n_routes + 1) * nRch, 0.0);
        std::vector<double> accB(wantBas ? nHRU : 0, 0.0);
        std::vector<double> stepX(anyStep ? (size_t)(o.n_routes + 1) * nRch : 0), accX(stepX.size(), 0.0), volNow(anyVol ? nRch : 0);
        std::vector<double> gageBuf(gageIdx.size());
        auto put_seg = [&](int var, size_t rec, const double *data) {      // a reach-dimensioned record, all reaches or the gauged ones
            if (!atGage) { w->put_record(var, rec, data); return; }
            for (size_t i = 0; i < gageIdx.size(); ++i) gageBuf[i] = data[gageIdx[i]];
            w->put_record(var, rec, gageBuf.data()); };
        auto put_volumes = [&](size_t rec) {                       // REACH_VOL(1) as the last step of the call left it
            for (int r = 0; r < o.n_routes; ++r) if (vVol[r] >= 0) {
                int e = mr_get_flux(h, o.route_methods[r], MR_REACH_VOL1, volNow.data(), msg); if (e) die(e, msg);
                put_seg(vVol[r], rec, volNow.data());
            }
        };
        int nAcc = 0; size_t recOut = 0, fileNo = 0; double tAcc = 0.0;
        // --device-history: the period means of the discharges and of dlayRunoff are formed on the device (mr_history_means,
        // histVars_data.f90:154-246) and only they travel to the host -- one record per output period instead of one per step.
        // Taken when the output is aggregated and asks for nothing the library would have to hand over step by step.
        bool devHist = deviceHistory && nAgg > 1 && !anyStep && !anyVol;
        if (deviceHistory && !devHist) std::fprintf(stderr, "route_runoff: --device-history needs <outputFrequency> > 1 and no per-step / volume variables; the host aggregates\n");
        const int nSer = o.n_routes + (wantDlay ? 1 : 0);
        std::vector<float> hist(devHist ? (size_t)(batch / nAgg + 2) * nSer * nRch : 0);
        std::vector<double> histRow(devHist ? nRch : 0);
        int histPer = 0;
        // the open output period goes through a restart as the reference's histVars does (nt, history_time, the running sums under
        // their history-file names); times are stored as absolute seconds so that a continuation run with a later <sim_start> reads them
        HistState hs;
        if (nAgg > 1) {
            for (int r = 0; r < o.n_routes; ++r) if (c.flag(vname[o.route_methods[r]], true)) hs.vars.push_back({vname[o.route_methods[r]], &acc, (size_t)r * nRch, nRch, false});
            if (wantDlay) hs.vars.push_back({"dlayRunoff", &acc, (size_t)o.n_routes * nRch, nRch, false});
            for (int r = 0; r < o.n_routes; ++r) if (wantInf[r]) hs.vars.push_back({infName[o.route_methods[r]], &accX, (size_t)r * nRch, nRch, false});
            if (wantInst) hs.vars.push_back({"instRunoff", &accX, (size_t)o.n_routes * nRch, nRch, false});
            if (wantBas) hs.vars.push_back({"basRunoff", &accB, 0, nHRU, true});
        }
        double T0 = 0.0;                                                       // TSEC(1) of a cold start, init_model_data.f90:600
        const std::string stateIn = c.str("fname_state_in", "coldstart");
        if (!stateIn.empty() && lower(stateIn) != "coldstart" && stateIn != "INPUT_RESTART_NC")
            T0 = read_restart(h, join_path(c.str("restart_dir", outdir), stateIn), o, segId, qmodOption == 1, nAgg > 1 ? &hs : nullptr);   // init_state_data, init_model_data.f90:332-623
            if (hs.nt > 0) { nAcc = hs.nt; tAcc = hs.tb[0] - tStartAsked; }      // the restart was written inside an output period
        if (o.is_lake_sim) {                                  // the library counts steps from the cold start: step 0 was T0 seconds before <sim_start>
            const Civil cv = civil_from_sec(tStart - T0, noleap);
            ierr = mr_set_sim_start(h, cv.y, cv.mo, cv.d, (double)cv.sod, noleap ? 1 : 0, msg); if (ierr) die(ierr, msg);
        }
        if (devHist) {
            bool inside = nAcc > 0;
            for (const auto &rp : restartPlan) if ((rp.first + 1 + (size_t)nAcc) % (size_t)nAgg != 0 && rp.first + 1 != nSteps) inside = true;
            if (inside) { devHist = false; std::fprintf(stderr, "route_runoff: a restart file falls inside an output period; the host aggregates (--device-history ignored)\n"); }
        }
        size_t nextRestart = 0;
        for (size_t s = 0; s < nSteps;) {
            int nb = (int)std::min<size_t>(batch, nSteps - s);
            if (nextRestart < restartPlan.size()) nb = (int)std::min<size_t>(nb, restartPlan[nextRestart].first + 1 - s);      // a batch ends where a restart file is due
            if (anyStep) nb = 1;
            else if (anyVol) nb = std::min(nb, nAgg - nAcc);                   // the batch ends where the output period ends
            if (deviceIngest && !wantBas) ingest_batch(s, nb);
            else for (int k = 0; k < nb; ++k) load_step(s + k, &ro[(size_t)k * inCols]);
            if (lakeForcing) {
                for (int k = 0; k < nb; ++k) { load_var(vsEvapo, s + k, &evRows[(size_t)k * nHRU], nHRU, false); load_var(vsPrecip, s + k, &prRows[(size_t)k * nHRU], nHRU, false); }
                ierr = mr_upload_lake_forcing(h, nb, evRows.data(), prRows.data(), msg); if (ierr) die(ierr, msg);
            }
            if (fluxWm || volWm) {
                for (int k = 0; k < nb; ++k) {
                    if (fluxWm) load_wm(c.need("vname_flux_wm"), false, s + k, &wmFluxRows[(size_t)k * nRch]);
                    if (volWm) load_wm(c.need("vname_vol_wm"), true, s + k, &wmVolRows[(size_t)k * nRch]);
                }
                ierr = mr_upload_wm(h, nb, fluxWm ? wmFluxRows.data() : nullptr, volWm ? wmVolRows.data() : nullptr, c.flag("is_vol_wm_jumpstart", false) ? 1 : 0, msg);
                if (ierr) die(ierr, msg);
            }
            if (qmodOption == 1) {
                for (int k = 0; k < nb; ++k) obsHas[k] = load_obs(tStart + (double)(s + k) * o.dt, &obsRows[(size_t)k * nRch]) ? 1 : 0;
                ierr = mr_upload_obs(h, nb, obsHas.data(), obsRows.data(), msg); if (ierr) die(ierr, msg);
            }
            if (deviceIngest) {
                if (wantBas) ingest_batch(s, nb);                              // <basRunoff> wants the rows on the host as well
                ierr = mr_route_resident(h, nb, T0, msg); if (ierr) die(ierr, msg);
                if (!devHist) { ierr = mr_download_q(h, nb, q.data(), msg); if (ierr) die(ierr, msg); }
            } else {
                ierr = mr_step_batch(h, nb, T0, ro.data(), devHist ? nullptr : q.data(), msg); if (ierr) die(ierr, msg);
            }
            if (devHist) {
                int nPer = 0;
                ierr = mr_history_means(h, nb, nAgg, wantDlay ? 1 : 0, s + nb == nSteps ? 1 : 0, (int)(hist.size() / ((size_t)nSer * nRch)), hist.data(), &nPer, msg);
                if (ierr) die(ierr, msg);
                histPer = 0;
            } else if (wantDlay) { ierr = mr_download_basin_q(h, nb, qd.data(), msg); if (ierr) die(ierr, msg); }
            if (anyStep) {                                                     // nb == 1: REACH_INFLOW / BASIN_QI of this step
                for (int r = 0; r < o.n_routes; ++r) if (wantInf[r]) { ierr = mr_get_flux(h, o.route_methods[r], MR_REACH_INFLOW, &stepX[(size_t)r * nRch], msg); if (ierr) die(ierr, msg); }
                if (wantInst) { ierr = mr_get_flux(h, o.route_methods[0], MR_BASIN_QI, &stepX[(size_t)o.n_routes * nRch], msg); if (ierr) die(ierr, msg); }
            }
            for (int k = 0; k < nb; ++k) {
                const double tsec = (tStart - tStartAsked) + (double)(s + k) * o.dt;    // seconds since <sim_start>
                if (fileNo < plan.size() && plan[fileNo].first == s + k) { open_history(plan[fileNo++].path); recOut = 0; }     // main_new_file
                if (nAgg == 1) {
                    { const double ts = tsec + stampOffset, tb[2] = {tsec, tsec + o.dt}; w->put_record(vTime, recOut, &ts); w->put_record(vTb, recOut, tb); }
                    for (int r = 0; r < o.n_routes; ++r) if (vQ[r] >= 0) put_seg(vQ[r], recOut, &q[((size_t)r * nb + k) * nRch]);
                    if (vDlay >= 0) put_seg(vDlay, recOut, &qd[(size_t)k * nRch]);
                    for (int r = 0; r < o.n_routes; ++r) if (vInf[r] >= 0) put_seg(vInf[r], recOut, &stepX[(size_t)r * nRch]);
                    if (vInst >= 0) put_seg(vInst, recOut, &stepX[(size_t)o.n_routes * nRch]);
                    if (anyVol) put_volumes(recOut);
                    if (vBas >= 0) w->put_record(vBas, recOut, &ro[(size_t)k * inCols]);
                    ++recOut;
                    continue;
                }
                if (nAcc == 0) { tAcc = tsec; std::fill(acc.begin(), acc.end(), 0.0); std::fill(accX.begin(), accX.end(), 0.0); }
                if (nAcc == 0) std::fill(accB.begin(), accB.end(), 0.0);
                for (size_t i = 0; i < accX.size(); ++i) accX[i] += stepX[i];
                for (size_t i = 0; i < accB.size(); ++i) accB[i] += ro[(size_t)k * inCols + i];
                if (!devHist) {
                    for (int r = 0; r < o.n_routes; ++r) for (size_t i = 0; i < nRch; ++i) acc[(size_t)r * nRch + i] += q[((size_t)r * nb + k) * nRch + i];
                    if (vDlay >= 0) for (size_t i = 0; i < nRch; ++i) acc[(size_t)o.n_routes * nRch + i] += qd[(size_t)k * nRch + i];
                }
                if (++nAcc == nAgg || s + k + 1 == nSteps) {
                    if (devHist) {                      // the next period of this batch, as the device formed it (float32 values)
                        for (int r = 0; r < nSer; ++r) { const float *src = &hist[((size_t)histPer * nSer + r) * nRch]; double *dst = &acc[(size_t)(r < o.n_routes ? r : o.n_routes) * nRch]; for (size_t i = 0; i < nRch; ++i) dst[i] = (double)src[i]; }
                        ++histPer;
                    } else
                    for (auto &v : acc) v /= (double)nAcc;
                    { const double ts = tAcc + stampOffset, tb[2] = {tAcc, tsec + o.dt}; w->put_record(vTime, recOut, &ts); w->put_record(vTb, recOut, tb); }
                    for (int r = 0; r < o.n_routes; ++r) if (vQ[r] >= 0) put_seg(vQ[r], recOut, &acc[(size_t)r * nRch]);
                    if (vDlay >= 0) put_seg(vDlay, recOut, &acc[(size_t)o.n_routes * nRch]);
                    for (auto &v : accX) v /= (double)nAcc;
                    for (int r = 0; r < o.n_routes; ++r) if (vInf[r] >= 0) put_seg(vInf[r], recOut, &accX[(size_t)r * nRch]);
                    if (vInst >= 0) put_seg(vInst, recOut, &accX[(size_t)o.n_routes * nRch]);
                    if (anyVol) put_volumes(recOut);
                    if (vBas >= 0) { for (auto &v : accB) v /= (double)nAcc; w->put_record(vBas, recOut, accB.data()); }
                    ++recOut; nAcc = 0;
                }
            }
            T0 += nb * o.dt; s += nb;
            if (nextRestart < restartPlan.size() && restartPlan[nextRestart].first + 1 == s) {                              // main_restart, route_runoff.f90:102
                hs.nt = nAgg > 1 ? nAcc : 0; hs.tb[0] = tStartAsked + tAcc; hs.tb[1] = tStartAsked + (tStart - tStartAsked) + (double)s * o.dt;
                write_restart(h, restartPlan[nextRestart].second, o, segId, T0, (long)std::lround(T0 / o.dt), qmodOption == 1, nAgg > 1 ? &hs : nullptr);
                std::printf("{\"restart\": \"%s\"}\n", restartPlan[nextRestart].second.c_str());
                ++nextRestart;
            }
        }
        for (auto *p : rd) delete p;
        if (w) w->close();
        mr_destroy(h);
        std::printf("{\"history\": \"%s\", \"steps\": %zu, \"history_files\": [", plan[0].path.c_str(), nSteps);
        for (size_t i = 0; i < plan.size(); ++i) std::printf("%s\"%s\"", i ? ", " : "", plan[i].path.c_str());
        std::printf("]}\n");
    } catch (const std::exception &e) {
        die(20, e.what());
    }
    return 0;
}
