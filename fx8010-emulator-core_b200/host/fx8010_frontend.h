// fx8010_frontend.h — host-side front-end of the B200 FX8010 executor: `.da` source text ->
// decoded program image (register table, instruction list, TRAM sizes, controls, metadata,
// error list) plus the host-built LOG/EXP tables.
//
// It accepts and rejects exactly what the reference's loader does (reference
// source/FX8010.cpp:777-875 loadFile, :365-741 syntaxCheck, :745-774 mapRegisterToIndex), but is
// written as a hand-rolled line scanner instead of seven std::regex objects per line; the
// differential tests in tests/test_frontend.py hold it against the compiled reference.
#pragma once

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "fx8010_gpu.h"

namespace fx8010 {

struct Register {                 // reference struct GPR, include/FX8010.h:167-174
    int type = FX_REG_STATIC;
    std::string name;
    float value = 0.0f;
    int io_index = 0;
};

struct Diagnostic {               // reference struct MyError, include/FX8010.h:63-67
    std::string description;
    int row = 1;
};

enum ErrorCode {                  // reference enum ErrorCode, include/FX8010.h:254-268
    ERR_NONE = 0, ERR_INVALID_INPUT, ERR_DIVISION_BY_ZERO, ERR_MULTIPLE_VAR_DECLARE, ERR_VAR_NOT_DECLARED,
    ERR_INPUT_FOR_R_NOT_ALLOWED, ERR_NO_END_FOUND, ERR_IO_INDEX_OUT_OF_RANGE, ERR_SYNTAX_NOT_VALID,
    ERR_ITRAMSIZE_TOO_LARGE, ERR_XTRAMSIZE_TOO_LARGE
};

constexpr int kMaxIDelay = 8192;      // MAX_IDELAY_SIZE, include/FX8010.h:41
constexpr int kMaxXDelay = 1048576;   // MAX_XDELAY_SIZE, include/FX8010.h:42

class Frontend {
public:
    explicit Frontend(int num_channels);
    void initialize();

    // One source file / text.  Appends to what earlier loads left behind, as the reference does.
    // Returns true when the error list holds nothing but the leading "no error" entry.
    bool loadFile(const std::string& path);
    bool loadText(const std::string& text);
    // One already comment-stripped, lower-cased line; the row used for diagnostics is row_counter.
    bool parseLine(const std::string& line);

    const std::vector<Register>& registers() const { return registers_; }
    std::vector<Register>& registers() { return registers_; }
    const std::vector<fx8010_instr>& instructions() const { return instructions_; }
    const std::vector<std::string>& controls() const { return controls_; }
    const std::unordered_map<std::string, std::string>& metadata() const { return meta_; }
    const std::vector<Diagnostic>& errors() const { return errors_; }
    int itramSize() const { return itram_size_; }
    int xtramSize() const { return xtram_size_; }
    int channels() const { return num_channels_; }
    void setChannels(int c) { num_channels_ = c; }
    bool ready() const { return ready_; }
    unsigned long generation() const { return generation_; }   // bumped by every initialize() / loadText(): the decoded image may have changed
    // Relaxed mode (off by default = the reference's exact accept/reject behaviour): also accepts what the
    // reference's README shows but its patterns reject: `itramsize N` with any or no
    // trailing blanks, CR-LF line ends, blanks around and blank lines after the final `end`.
    void setRelaxed(bool on) { relaxed_ = on; }
    bool relaxed() const { return relaxed_; }
    int findRegister(const std::string& name) const;     // -1 when absent (first match wins)

    // LOG / EXP tables, [32][64] doubles each (reference source/FX8010.cpp:63-105, 129-199).
    const std::vector<double>& logTables() const { return log_tables_; }
    const std::vector<double>& expTables() const { return exp_tables_; }

    // C-ABI view of the decoded program; pointers stay valid until the next load / parse.
    const fx8010_program_image* image();

    static std::string errorText(int code, int num_channels);

private:
    int mapOperand(const std::string& token);
    void report(int code);
    bool parseDeclaration(const std::string& s);
    bool parseTramSize(const std::string& s, bool& matched);
    bool parseInstruction(const std::string& s, bool& matched);
    void buildTables();

    int num_channels_;
    int channels_at_init_ = 1;        // the I/O-range message is fixed when the object is built (:32)
    std::vector<Register> registers_;
    std::vector<fx8010_instr> instructions_;
    std::vector<std::string> controls_;
    std::unordered_map<std::string, std::string> meta_;
    std::vector<Diagnostic> errors_;
    int row_counter_ = 1;             // reference errorCounter: never reset between loads
    int itram_size_ = 0, xtram_size_ = 0;
    bool ready_ = false;
    unsigned long generation_ = 0;
    bool relaxed_ = false;
    std::vector<double> log_tables_, exp_tables_;
    std::vector<fx8010_reg> image_regs_;
    fx8010_program_image image_ = {};
};

// `^-?\d+(\.\d+)?$` — decides whether an undeclared operand becomes a literal register
// (reference source/helpers.cpp:21-27).
bool isNumber(const std::string& s);

}  // namespace fx8010
