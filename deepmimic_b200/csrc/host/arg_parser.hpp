// `--key v1 v2 ...` argument files / argv lists, same observable semantics as the
// reference parser (R/DeepMimicCore/util/ArgParser.cpp:32-120, 258-300):
//   * a token is a key iff it has >= 3 chars and starts with "--"
//   * tokens (argv) or whole lines (files) starting with '#' are comments
//   * file tokens split on space, tab, CR, LF and ','
//   * the FIRST occurrence of a key wins (command line is loaded before --arg_file, so the
//     command line overrides the file: R/DeepMimicCore/DeepMimicCore.cpp:25-44)
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace dmh {

class ArgParser {
public:
    void LoadArgs(const std::vector<std::string>& toks) {
        std::vector<std::string> vals;
        std::string key;
        auto flush = [&]() {
            if (!key.empty() && table_.find(key) == table_.end()) table_[key] = vals;
        };
        for (const std::string& t : toks) {
            if (!t.empty() && t[0] == '#') continue;
            if (t.size() >= 3 && t[0] == '-' && t[1] == '-') {
                flush();
                vals.clear();
                key = t.substr(2);
            } else {
                vals.push_back(t);
            }
        }
        flush();
    }

    bool LoadFile(const std::string& path) {
        std::ifstream f(path);
        if (!f.good()) return false;
        std::vector<std::string> toks;
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::string buf;
            for (char c : line) {
                if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == ',') {
                    if (!buf.empty()) { toks.push_back(buf); buf.clear(); }
                } else {
                    buf += c;
                }
            }
            if (!buf.empty()) toks.push_back(buf);
        }
        LoadArgs(toks);
        return true;
    }

    bool Has(const std::string& k) const { return table_.find(k) != table_.end(); }

    bool ParseString(const std::string& k, std::string& out) const {
        auto it = table_.find(k);
        if (it == table_.end() || it->second.empty()) return false;
        out = it->second[0];
        return true;
    }
    bool ParseStrings(const std::string& k, std::vector<std::string>& out) const {
        auto it = table_.find(k);
        if (it == table_.end()) return false;
        out = it->second;
        return true;
    }
    bool ParseInt(const std::string& k, int& out) const {
        std::string s;
        if (!ParseString(k, s)) return false;
        out = std::atoi(s.c_str());
        return true;
    }
    bool ParseInts(const std::string& k, std::vector<int>& out) const {
        auto it = table_.find(k);
        if (it == table_.end()) return false;
        out.clear();
        for (const auto& s : it->second) out.push_back(std::atoi(s.c_str()));
        return true;
    }
    bool ParseDouble(const std::string& k, double& out) const {
        std::string s;
        if (!ParseString(k, s)) return false;
        out = std::atof(s.c_str());
        return true;
    }
    bool ParseDoubles(const std::string& k, std::vector<double>& out) const {
        auto it = table_.find(k);
        if (it == table_.end()) return false;
        out.clear();
        for (const auto& s : it->second) out.push_back(std::atof(s.c_str()));
        return true;
    }
    bool ParseBool(const std::string& k, bool& out) const {
        std::string s;
        if (!ParseString(k, s)) return false;
        out = (s == "true" || s == "1" || s == "True" || s == "T" || s == "t");
        return true;
    }

private:
    std::map<std::string, std::vector<std::string>> table_;
};

}  // namespace dmh
