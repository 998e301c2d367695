// fx8010_frontend.cpp — see fx8010_frontend.h.  Host-only; no CUDA here.
#include "fx8010_frontend.h"

#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace fx8010 {

namespace {

// Character classes of the reference's ECMAScript patterns in the "C" locale.
inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_word(char c) { return is_digit(c) || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
inline bool is_operand_char(char c) { return is_word(c) || c == '.' || c == '-'; }   // [a-zA-Z0-9_.-]

size_t skip_space(const std::string& s, size_t i) {
    while (i < s.size() && is_space(s[i])) ++i;
    return i;
}
bool all_space(const std::string& s, size_t i) { return skip_space(s, i) == s.size(); }

// The first blank-delimited word of the line and where it ends.  Every keyword of the dialect
// must be followed by whitespace (or, for `end`, by nothing), so a keyword matches exactly when
// this word equals it.
std::string first_word(const std::string& s, size_t& end) {
    const size_t b = skip_space(s, 0);
    size_t e = b;
    while (e < s.size() && !is_space(s[e])) ++e;
    end = e;
    return s.substr(b, e - b);
}

// std::stoi on a string of digits; false when it would throw (empty or out of int range).
bool digits_to_int(const std::string& d, int& out) {
    if (d.empty()) return false;
    errno = 0;
    const long v = std::strtol(d.c_str(), nullptr, 10);
    if (errno == ERANGE || v > INT_MAX || v < INT_MIN) return false;
    out = (int)v;
    return true;
}

const char* const kDeclWords[] = {"static", "temp", "control", "input", "output", "const"};
const int kDeclTypes[] = {FX_REG_STATIC, FX_REG_TEMP, FX_REG_CONTROL, FX_REG_INPUT, FX_REG_OUTPUT, FX_REG_CONST};

struct OpWord { const char* word; int opcode; };
const OpWord kOpWords[] = {                       // reference opcodeMap, include/FX8010.h:103-122
    {"macs", FX_MACS}, {"macsn", FX_MACSN}, {"macints", FX_MACINTS}, {"macintw", FX_MACINTW}, {"acc3", FX_ACC3},
    {"macmv", FX_MACMV}, {"macw", FX_MACW}, {"macwn", FX_MACWN}, {"skip", FX_SKIP}, {"andxor", FX_ANDXOR},
    {"tstneg", FX_TSTNEG}, {"limit", FX_LIMIT}, {"limitn", FX_LIMITN}, {"log", FX_LOG}, {"exp", FX_EXP},
    {"interp", FX_INTERP}, {"idelay", FX_IDELAY}, {"xdelay", FX_XDELAY}};
const char* const kMetaWords[] = {"name", "copyright", "created", "engine", "comment", "guid"};

}  // namespace

bool isNumber(const std::string& s) {
    size_t i = 0;
    if (i < s.size() && s[i] == '-') ++i;
    size_t d = i;
    while (i < s.size() && is_digit(s[i])) ++i;
    if (i == d) return false;
    if (i == s.size()) return true;
    if (s[i] != '.') return false;
    d = ++i;
    while (i < s.size() && is_digit(s[i])) ++i;
    return i > d && i == s.size();
}

std::string Frontend::errorText(int code, int num_channels) {
    switch (code) {                               // reference source/FX8010.cpp:25-35
    case ERR_NONE: return "Kein Fehler";
    case ERR_INVALID_INPUT: return "Ungueltige Eingabe";
    case ERR_DIVISION_BY_ZERO: return "Division durch Null";
    case ERR_MULTIPLE_VAR_DECLARE: return "Mehrfache Variablendeklaration";
    case ERR_VAR_NOT_DECLARED: return "Variable nicht deklariert";
    case ERR_INPUT_FOR_R_NOT_ALLOWED: return "Verwendung von Input fuer R ist nicht erlaubt";
    case ERR_NO_END_FOUND: return "Kein 'END' gefunden";
    case ERR_IO_INDEX_OUT_OF_RANGE: return "I/O Index ausserhalb des gueltigen Bereichs (max. " + std::to_string(num_channels) + ")";
    case ERR_SYNTAX_NOT_VALID: return "Ungueltige Syntax";
    case ERR_ITRAMSIZE_TOO_LARGE: return "iTRAM Size ausserhalb des gueltigen Bereichs (max. " + std::to_string(kMaxIDelay) + ")";
    case ERR_XTRAMSIZE_TOO_LARGE: return "xRAM Size ausserhalb des gueltigen Bereichs (max. " + std::to_string(kMaxXDelay) + ")";
    default: return "";
    }
}

Frontend::Frontend(int num_channels) : num_channels_(num_channels) { initialize(); }

// What the reference's constructor-time initialize() sets up (source/FX8010.cpp:16-125).  Like
// there it is public and APPENDS when called again; lookups find registers 0..3 first either way.
void Frontend::initialize() {
    ++generation_;
    channels_at_init_ = num_channels_;
    errors_.clear();
    errors_.push_back({errorText(ERR_NONE, num_channels_), 1});   // entry 0 is always "no error" (:38-42)
    registers_.push_back({FX_REG_CCR, "ccr", 0.0f, 0});           // GPR 0..3 are reserved (:50-55)
    registers_.push_back({FX_REG_READ, "read", 0.0f, 0});
    registers_.push_back({FX_REG_WRITE, "write", 0.0f, 0});
    registers_.push_back({FX_REG_AT, "at", 0.0f, 0});
    buildTables();
}

// Tables for exponent e: 32 points of x^(1/e) (LOG) or x^e (EXP) on [0,1] form the upper half; the
// lower half is their mirror image negated THROUGH A FLOAT, as the reference's range-for over a
// `float` loop variable does (source/FX8010.cpp:190-199), so it is float-rounded.
void Frontend::buildTables() {
    const int half = FX8010_TABLE_ENTRIES / 2;
    log_tables_.assign((size_t)FX8010_TABLE_COUNT * FX8010_TABLE_ENTRIES, 0.0);
    exp_tables_.assign((size_t)FX8010_TABLE_COUNT * FX8010_TABLE_ENTRIES, 0.0);
    const double x_min = 0, x_max = 1.0;
    const double step = (x_max - x_min) / (half - 1);
    for (int e = 0; e < FX8010_TABLE_COUNT; ++e) {
        double* L = &log_tables_[(size_t)e * FX8010_TABLE_ENTRIES];
        double* X = &exp_tables_[(size_t)e * FX8010_TABLE_ENTRIES];
        for (int i = 0; i < half; ++i) {
            const double x = x_min + (i * step);
            L[half + i] = std::pow(x, 1.0 / static_cast<float>(e));
            X[half + i] = std::pow(x, static_cast<float>(e));
        }
        for (int i = 0; i < half; ++i) {
            const float fl = static_cast<float>(L[FX8010_TABLE_ENTRIES - 1 - i]);
            const float fx = static_cast<float>(X[FX8010_TABLE_ENTRIES - 1 - i]);
            L[i] = -fl;
            X[i] = -fx;
        }
    }
}

int Frontend::findRegister(const std::string& name) const {
    for (size_t i = 0; i < registers_.size(); ++i)
        if (registers_[i].name == name) return (int)i;
    return -1;
}

void Frontend::report(int code) {
    errors_.push_back({code == ERR_IO_INDEX_OUT_OF_RANGE ? errorText(code, channels_at_init_) : errorText(code, num_channels_), row_counter_});
}

// Undeclared numeric operands become STATIC literal registers named by their spelling
// (reference source/FX8010.cpp:745-774): "0", "0.0", "1" and "1.0" are four different registers.
int Frontend::mapOperand(const std::string& token) {
    const int idx = findRegister(token);
    if (idx >= 0) return idx;
    if (!isNumber(token)) return -1;
    Register r;
    r.type = FX_REG_STATIC;
    r.name = token;
    r.value = std::strtof(token.c_str(), nullptr);
    registers_.push_back(r);
    return (int)registers_.size() - 1;
}

// (static|temp|control|input|output|const) name [ [\s=,]* number ]      (source/FX8010.cpp:371, 395-485)
// The reference pattern backtracks over the length of `name`: "static a1.5" declares `a` = 1.5.
bool Frontend::parseDeclaration(const std::string& s) {
    size_t kw_end;
    const std::string word = first_word(s, kw_end);
    int type = -1;
    for (int k = 0; k < 6; ++k) if (word == kDeclWords[k]) type = kDeclTypes[k];
    if (type < 0 || kw_end >= s.size()) return false;
    const size_t nb = skip_space(s, kw_end);
    size_t ne = nb;
    while (ne < s.size() && is_word(s[ne])) ++ne;
    std::string name, value;
    bool matched = false;
    for (size_t len = ne - nb; len >= 1 && !matched; --len) {
        const size_t k = nb + len;
        if (all_space(s, k)) { name = s.substr(nb, len); value.clear(); matched = true; break; }
        size_t m = k;
        while (m < s.size() && (is_space(s[m]) || s[m] == '=' || s[m] == ',')) ++m;
        size_t d = m;
        while (d < s.size() && is_digit(s[d])) ++d;
        if (d == m) continue;
        if (d + 1 < s.size() && s[d] == '.' && is_digit(s[d + 1])) {
            size_t f = d + 1;
            while (f < s.size() && is_digit(s[f])) ++f;
            d = f;
        }
        if (!all_space(s, d)) continue;
        name = s.substr(nb, len); value = s.substr(m, d - m); matched = true;
    }
    if (!matched) return false;

    // from here on the line IS a declaration; the result is what the reference's handler returns
    if (type == FX_REG_CONTROL) controls_.push_back(name);                // listed before the duplicate check (:408-411)
    if (findRegister(name) != -1) { report(ERR_MULTIPLE_VAR_DECLARE); return true; }
    Register reg;
    reg.type = type;
    reg.name = name;
    if (!value.empty()) {
        if (type == FX_REG_INPUT || type == FX_REG_OUTPUT) {              // the number is the channel (:443-459)
            int io = 0;
            const std::string int_part = value.substr(0, value.find('.'));
            if (!digits_to_int(int_part, io) || io > num_channels_ - 1) { report(ERR_IO_INDEX_OUT_OF_RANGE); return true; }
            reg.io_index = io;
        } else {
            reg.value = std::strtof(value.c_str(), nullptr);              // :463
        }
    }
    registers_.push_back(reg);
    return true;
}

// (itramsize|xtramsize)\s+(\d+)*\s$ — exactly ONE blank after the number (source/FX8010.cpp:377, 498-543)
bool Frontend::parseTramSize(const std::string& s, bool& matched) {
    matched = false;
    size_t kw_end;
    const std::string word = first_word(s, kw_end);
    const bool is_i = (word == "itramsize"), is_x = (word == "xtramsize");
    if (!is_i && !is_x) return false;
    const size_t q = skip_space(s, kw_end);
    if (q == kw_end) return false;                                        // needs \s+
    size_t r = q;
    while (r < s.size() && is_digit(s[r])) ++r;
    std::string digits;
    if (r > q) {
        if (relaxed_ ? !all_space(s, r) : !(r + 1 == s.size() && is_space(s[r]))) return false;
        digits = s.substr(q, r - q);
    } else {
        if (!(q == s.size() && q - kw_end >= 2)) return false;            // "itramsize  ": matches with an empty number
    }
    matched = true;
    int size = 0;
    // The reference calls std::stoi on the capture; an empty or oversized capture would throw
    // there (process abort).  Here that line is reported as a syntax error instead.
    const bool convertible = digits_to_int(digits, size);
    if (is_i) {
        if (itram_size_ > kMaxIDelay) { report(ERR_ITRAMSIZE_TOO_LARGE); return true; }   // tests the PREVIOUS size (:506)
        if (!convertible) { report(ERR_SYNTAX_NOT_VALID); return true; }
        itram_size_ = size;
    } else {
        if (xtram_size_ > kMaxXDelay) { report(ERR_XTRAMSIZE_TOO_LARGE); return true; }
        if (!convertible) { report(ERR_SYNTAX_NOT_VALID); return true; }
        xtram_size_ = size;
    }
    return true;
}

// op R, A, X, Y   (source/FX8010.cpp:380, 548-695)
bool Frontend::parseInstruction(const std::string& s, bool& matched) {
    matched = false;
    size_t kw_end;
    const std::string word = first_word(s, kw_end);
    int opcode = -1;
    for (const OpWord& w : kOpWords) if (word == w.word) opcode = w.opcode;
    if (opcode < 0 || kw_end >= s.size()) return false;
    std::string ops[4];
    size_t i = skip_space(s, kw_end);
    for (int k = 0; k < 4; ++k) {
        size_t e = i;
        while (e < s.size() && is_operand_char(s[e])) ++e;
        if (e == i) return false;
        ops[k] = s.substr(i, e - i);
        i = skip_space(s, e);
        if (k < 3) {
            if (i >= s.size() || s[i] != ',') return false;
            i = skip_space(s, i + 1);
        }
    }
    if (i != s.size()) return false;
    matched = true;

    fx8010_instr ins = {};
    ins.opcode = opcode;
    int idx[4];
    for (int k = 0; k < 4; ++k) {
        idx[k] = mapOperand(ops[k]);      // literal registers created so far stay even if a later operand fails
        if (idx[k] < 0) { report(ERR_VAR_NOT_DECLARED); return true; }
        const Register& reg = registers_[idx[k]];
        if (k == 0) {
            if (reg.type == FX_REG_INPUT) { report(ERR_INPUT_FOR_R_NOT_ALLOWED); return true; }
            if (reg.type == FX_REG_OUTPUT) ins.has_output = 1;
        } else {
            if (reg.type == FX_REG_INPUT) ins.has_input = 1;
            else if (reg.name == "noise") ins.has_noise = 1;
        }
    }
    ins.r = idx[0]; ins.a = idx[1]; ins.x = idx[2]; ins.y = idx[3];
    instructions_.push_back(ins);
    return true;
}

bool Frontend::parseLine(const std::string& s) {
    const size_t before = errors_.size();
    bool matched = false;
    if (parseDeclaration(s)) return errors_.size() == before;
    if (all_space(s, 0)) return true;                                     // blank (:374)
    if (parseTramSize(s, matched), matched) return errors_.size() == before;
    if (parseInstruction(s, matched), matched) return errors_.size() == before;
    size_t kw_end;
    const std::string word = first_word(s, kw_end);
    for (const char* m : kMetaWords) {                                    // key "value"   (:383, 699-708)
        if (word != m) continue;
        const size_t q = skip_space(s, kw_end);
        if (q == kw_end || q >= s.size() || s[q] != '"') break;
        const size_t close = s.find('"', q + 1);
        if (close == std::string::npos || close == q + 1 || close + 1 != s.size()) break;
        meta_[word] = s.substr(q + 1, close - q - 1);
        return true;
    }
    if (word == "end" && all_space(s, kw_end)) {                          // :386, 712-718
        fx8010_instr end = {};
        end.opcode = FX_END;
        instructions_.push_back(end);
        return true;
    }
    {                                                                     // ";"-only line (:389); unreachable after comment stripping
        size_t i = skip_space(s, 0), semi = i;
        while (semi < s.size() && s[semi] == ';') ++semi;
        if (semi > i && all_space(s, semi)) return true;
    }
    report(ERR_SYNTAX_NOT_VALID);
    return false;
}

bool Frontend::loadText(const std::string& text) {
    ++generation_;
    std::istringstream in(text);
    std::vector<std::string> lines;
    std::string line;
    while (std::getline(in, line)) {
        const size_t c = line.find(';');                                  // comments run to the end of the line (:794-798)
        if (c != std::string::npos) line.resize(c);
        for (char& ch : line) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');   // :805-808
        if (relaxed_ && !line.empty() && line.back() == '\r') line.pop_back();
        lines.push_back(line);
    }
    if (relaxed_) {                                                       // blank lines after / blanks around the final `end`
        while (!lines.empty() && all_space(lines.back(), 0)) lines.pop_back();
        if (!lines.empty()) {
            size_t e;
            if (first_word(lines.back(), e) == "end" && all_space(lines.back(), e)) lines.back() = "end";
        }
    }
    for (const std::string& l : lines) { parseLine(l); ++row_counter_; }  // errors do not stop the scan (:819-825)
    if (lines.empty() || lines.back() != "end") report(ERR_NO_END_FOUND); // exact match: no blanks, no CR (:829-838)
    if (errors_.size() > 1) return false;
    ready_ = true;
    return true;
}

bool Frontend::loadFile(const std::string& path) {
    std::ifstream file(path, std::ios::binary);
    if (!file) return false;                                              // no error entry (:868-873)
    std::ostringstream ss;
    ss << file.rdbuf();
    return loadText(ss.str());
}

const fx8010_program_image* Frontend::image() {
    image_regs_.resize(registers_.size());
    for (size_t i = 0; i < registers_.size(); ++i) {
        image_regs_[i].type = registers_[i].type;
        image_regs_[i].init_value = registers_[i].value;
        image_regs_[i].io_index = registers_[i].io_index;
        image_regs_[i].is_noise = registers_[i].name == "noise";
    }
    image_.instrs = instructions_.data(); image_.n_instrs = (int)instructions_.size();
    image_.regs = image_regs_.data(); image_.n_regs = (int)image_regs_.size();
    image_.itram_size = itram_size_; image_.xtram_size = xtram_size_;
    image_.log_tables = log_tables_.data(); image_.exp_tables = exp_tables_.data();
    return &image_;
}

}  // namespace fx8010
