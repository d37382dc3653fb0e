#include "gpus_info.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <map>
#include <regex>

namespace {

// ---- a minimal JSON reader (objects, arrays, strings, numbers, true/false/null) -----------
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    long long inum = 0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal *get(const std::string &k) const {
        for (const auto &kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit JParser(const std::string &src) : s(src) {}
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    bool fail(const std::string &m) { if (err.empty()) err = m + " at offset " + std::to_string(i); return false; }
    bool lit(const char *w) {
        size_t n = std::char_traits<char>::length(w);
        if (s.compare(i, n, w) != 0) return fail(std::string("expected ") + w);
        i += n;
        return true;
    }
    bool str(std::string *out) {
        if (i >= s.size() || s[i] != '"') return fail("expected string");
        i++;
        out->clear();
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c == '\\') {
                if (i >= s.size()) return fail("bad escape");
                char e = s[i++];
                switch (e) {
                    case 'n': out->push_back('\n'); break;
                    case 't': out->push_back('\t'); break;
                    case 'r': out->push_back('\r'); break;
                    case 'b': out->push_back('\b'); break;
                    case 'f': out->push_back('\f'); break;
                    case 'u': {
                        if (i + 4 > s.size()) return fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(s.substr(i, 4).c_str(), nullptr, 16);
                        i += 4;
                        if (cp < 0x80) out->push_back((char)cp);
                        else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
                        else { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
                        break;
                    }
                    default: out->push_back(e);
                }
            } else {
                out->push_back(c);
            }
        }
        if (i >= s.size()) return fail("unterminated string");
        i++;
        return true;
    }
    bool value(JVal *v) {
        ws();
        if (i >= s.size()) return fail("unexpected end");
        char c = s[i];
        if (c == '{') {
            v->kind = JVal::Obj;
            i++;
            ws();
            if (i < s.size() && s[i] == '}') { i++; return true; }
            while (true) {
                ws();
                std::string k;
                if (!str(&k)) return false;
                ws();
                if (i >= s.size() || s[i] != ':') return fail("expected ':'");
                i++;
                JVal child;
                if (!value(&child)) return false;
                v->obj.emplace_back(std::move(k), std::move(child));
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == '}') { i++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v->kind = JVal::Arr;
            i++;
            ws();
            if (i < s.size() && s[i] == ']') { i++; return true; }
            while (true) {
                JVal child;
                if (!value(&child)) return false;
                v->arr.push_back(std::move(child));
                ws();
                if (i < s.size() && s[i] == ',') { i++; continue; }
                if (i < s.size() && s[i] == ']') { i++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v->kind = JVal::Str; return str(&v->str); }
        if (c == 't') { v->kind = JVal::Bool; v->b = true; return lit("true"); }
        if (c == 'f') { v->kind = JVal::Bool; v->b = false; return lit("false"); }
        if (c == 'n') { v->kind = JVal::Null; return lit("null"); }
        size_t j = i;
        if (j < s.size() && (s[j] == '-' || s[j] == '+')) j++;
        while (j < s.size() && (std::isdigit((unsigned char)s[j]) || s[j] == '.' || s[j] == 'e' || s[j] == 'E' || s[j] == '-' || s[j] == '+')) j++;
        if (j == i) return fail("unexpected character");
        const std::string tok = s.substr(i, j - i);
        v->kind = JVal::Num;
        v->num = strtod(tok.c_str(), nullptr);
        v->inum = strtoll(tok.c_str(), nullptr, 10);
        i = j;
        return true;
    }
};

std::string strOf(const JVal *v) { return v && v->kind == JVal::Str ? v->str : std::string(); }
long long intOf(const JVal *v) { return v && v->kind == JVal::Num ? v->inum : 0; }

}  // namespace

namespace nvgputypes {

std::string ParseGpusInfo(const std::string &json, GpusInfo *out) {
    JParser p(json);
    JVal root;
    if (!p.value(&root)) return "GpusInfo: " + p.err;
    p.ws();
    if (p.i != json.size()) return "GpusInfo: trailing characters at offset " + std::to_string(p.i);
    if (root.kind != JVal::Obj) return "GpusInfo: top level is not an object";
    *out = GpusInfo();
    if (const JVal *ver = root.get("Version")) {
        out->Driver = strOf(ver->get("Driver"));
        out->CUDA = strOf(ver->get("CUDA"));
    }
    const JVal *devs = root.get("Devices");
    if (!devs || devs->kind != JVal::Arr) return "";
    for (const JVal &d : devs->arr) {
        if (d.kind != JVal::Obj) return "GpusInfo: device entry is not an object";
        GpuInfo g;
        g.ID = strOf(d.get("UUID"));
        g.Model = strOf(d.get("Model"));
        g.Path = strOf(d.get("Path"));
        if (const JVal *mem = d.get("Memory")) g.MemoryGlobal = intOf(mem->get("Global"));
        if (const JVal *pci = d.get("PCI")) {
            g.BusID = strOf(pci->get("BusID"));
            g.Bandwidth = intOf(pci->get("Bandwidth"));
        }
        if (const JVal *topo = d.get("Topology"))
            if (topo->kind == JVal::Arr)
                for (const JVal &t : topo->arr) g.Topology.push_back({strOf(t.get("BusID")), (int32_t)intOf(t.get("Link"))});
        out->Gpus.push_back(std::move(g));
    }
    return "";
}

}  // namespace nvgputypes

namespace nvidia {

namespace {

void discoveryPass(nvgputypes::GpusInfo *info, const std::map<std::string, size_t> &busToIdx,
                   const std::vector<int32_t> &links, int level) {
    for (auto &g : info->Gpus) g.TopoDone = false;
    int linkID = 0;
    for (size_t i = 0; i < info->Gpus.size(); i++) {
        nvgputypes::GpuInfo &g = info->Gpus[i];
        if (!g.Found || g.TopoDone) continue;
        const std::string prefix = "gpugrp" + std::to_string(level) + "/" + std::to_string(linkID++);
        g.Name = prefix + "/" + g.Name;
        g.TopoDone = true;
        for (const auto &t : g.Topology) {
            if (std::find(links.begin(), links.end(), t.Link) == links.end()) continue;
            auto it = busToIdx.find(t.BusID);
            if (it == busToIdx.end()) continue;
            nvgputypes::GpuInfo &o = info->Gpus[it->second];
            if (o.Found) {                          // no TopoDone check in the reference (:80-87)
                o.Name = prefix + "/" + o.Name;
                o.TopoDone = true;
            }
        }
    }
}

}  // namespace

void DiscoverTopology(nvgputypes::GpusInfo *info, bool useNVML) {
    std::map<std::string, size_t> busToIdx;
    for (size_t i = 0; i < info->Gpus.size(); i++) {
        nvgputypes::GpuInfo &g = info->Gpus[i];
        if (!useNVML) {                              // nvidia_gpu_manager.go:124-129
            g.MemoryGlobal *= 1024LL * 1024LL;
            g.Bandwidth *= 1000LL * 1000LL;
        }
        g.Found = true;
        g.Index = (int)i;
        g.Name = "gpu/" + g.ID;
        busToIdx[g.BusID] = i;                       // later duplicates win, like the Go map
    }
    discoveryPass(info, busToIdx, {6, 5, 4}, 0);             // :178
    discoveryPass(info, busToIdx, {6, 5, 4, 3, 2, 1}, 1);    // :180
}

void UpdateNodeInfo(const nvgputypes::GpusInfo &named, types::NodeInfo *nodeInfo) {
    const int64_t n = (int64_t)named.Gpus.size();
    nodeInfo->Capacity[gpuplugintypes::ResourceGPU] = n;
    nodeInfo->Allocatable[gpuplugintypes::ResourceGPU] = n;
    nodeInfo->KubeCap[gpuplugintypes::ResourceGPU] = n;
    nodeInfo->KubeAlloc[gpuplugintypes::ResourceGPU] = n;
    for (const auto &g : named.Gpus) {
        if (!g.Found) continue;
        types::AddGroupResource(nodeInfo->Capacity, g.Name + "/memory", g.MemoryGlobal);
        types::AddGroupResource(nodeInfo->Allocatable, g.Name + "/memory", g.MemoryGlobal);
        types::AddGroupResource(nodeInfo->Capacity, g.Name + "/cards", 1);
        types::AddGroupResource(nodeInfo->Allocatable, g.Name + "/cards", 1);
    }
}

std::vector<int32_t> LinkMatrix(const nvgputypes::GpusInfo &info) {
    const size_t n = info.Gpus.size();
    std::map<std::string, size_t> busToIdx;
    for (size_t i = 0; i < n; i++) busToIdx.emplace(info.Gpus[i].BusID, i);
    std::vector<int32_t> m(n * n, 0);
    for (size_t i = 0; i < n; i++)
        for (const auto &t : info.Gpus[i].Topology) {
            auto it = busToIdx.find(t.BusID);
            if (it != busToIdx.end() && it->second != i) m[i * n + it->second] = t.Link;
        }
    return m;
}

std::string VisibleDevices(const types::ContainerInfo &cont) {
    static const std::regex rx(std::string(types::DeviceGroupPrefix) + "/gpugrp1/.*/gpugrp0/.*/gpu/(.*?)/cards");
    std::string out;
    std::smatch m;
    for (const auto &kv : cont.AllocateFrom)
        if (std::regex_search(kv.second, m, rx)) {
            if (!out.empty()) out += ",";
            out += m[1].str();
        }
    return out;
}

}  // namespace nvidia
