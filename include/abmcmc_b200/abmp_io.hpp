// ABMP: the flat, self-describing problem container of abmcmc_b200.
//
// The reference persists problems as Boost binary archives of its C++ object graph
// (PredPreyProblem.h:89-106, TableauNormMinimiser.h:150-156).  That format needs Boost and the
// reference's class layout.  ABMP is our own format: a list of named little-endian arrays that
// hold exactly what the SparseBasisSampler constructor consumes (SparseBasisSampler.h:107-185):
// the reduced tableau (non-basic columns over all tableau rows, L/U/F per row), the factor
// closures tabulated over each row's reachable domain, and the importance tables of the agent
// (TrajectoryImportance.h:90-113 evaluated through PredPreyAgent.h:124-182: pp_meta, imp_*; or, for any other agent type,
// through its own timestep / neighbours / consequences: ag_meta, ag_class, ag_infl, ag_nb*, ag_cq*, imp_* --
// reference_flatten.hpp).
//
// Layout:  "ABMPROB1" | u32 nArrays | nArrays x { char name[24] | u32 dtype | u64 count | data }
// dtype: 0 = int32, 1 = float64, 2 = uint8, 3 = int8.  Data is padded to 8 bytes.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace abmcmc {

struct AbmpFile {
    std::map<std::string, std::vector<int32_t>> i32;
    std::map<std::string, std::vector<double>> f64;
    std::map<std::string, std::vector<uint8_t>> u8;
    std::map<std::string, std::vector<int8_t>> i8;

    const std::vector<int32_t> &I(const std::string &k) const {
        auto it = i32.find(k);
        if (it == i32.end()) throw std::runtime_error("ABMP: missing int32 array '" + k + "'");
        return it->second;
    }
    const std::vector<double> &D(const std::string &k) const {
        auto it = f64.find(k);
        if (it == f64.end()) throw std::runtime_error("ABMP: missing float64 array '" + k + "'");
        return it->second;
    }
    const std::vector<int8_t> &B(const std::string &k) const {
        auto it = i8.find(k);
        if (it == i8.end()) throw std::runtime_error("ABMP: missing int8 array '" + k + "'");
        return it->second;
    }
    bool has(const std::string &k) const {
        return i32.count(k) || f64.count(k) || u8.count(k) || i8.count(k);
    }

    void write(const std::string &path) const {
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("ABMP: cannot open '" + path + "' for writing");
        std::fwrite("ABMPROB1", 1, 8, f);
        uint32_t n = uint32_t(i32.size() + f64.size() + u8.size() + i8.size());
        std::fwrite(&n, 4, 1, f);
        for (auto &kv : i32) put(f, kv.first, 0, kv.second.data(), kv.second.size(), 4);
        for (auto &kv : f64) put(f, kv.first, 1, kv.second.data(), kv.second.size(), 8);
        for (auto &kv : u8) put(f, kv.first, 2, kv.second.data(), kv.second.size(), 1);
        for (auto &kv : i8) put(f, kv.first, 3, kv.second.data(), kv.second.size(), 1);
        std::fclose(f);
    }

    static AbmpFile read(const std::string &path) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("ABMP: cannot open '" + path + "'");
        AbmpFile out;
        char magic[8];
        uint32_t n = 0;
        if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "ABMPROB1", 8) != 0 ||
            std::fread(&n, 4, 1, f) != 1) {
            std::fclose(f);
            throw std::runtime_error("ABMP: bad header in '" + path + "'");
        }
        for (uint32_t a = 0; a < n; ++a) {
            char name[24];
            uint32_t dtype;
            uint64_t count;
            if (std::fread(name, 1, 24, f) != 24 || std::fread(&dtype, 4, 1, f) != 1 ||
                std::fread(&count, 8, 1, f) != 1) {
                std::fclose(f);
                throw std::runtime_error("ABMP: truncated array header");
            }
            name[23] = 0;
            size_t esz = dtype == 0 ? 4 : dtype == 1 ? 8 : 1;
            size_t bytes = size_t(count) * esz, padded = (bytes + 7) & ~size_t(7);
            std::vector<uint8_t> buf(padded);
            if (padded && std::fread(buf.data(), 1, padded, f) != padded) {
                std::fclose(f);
                throw std::runtime_error("ABMP: truncated array data");
            }
            switch (dtype) {
                case 0: out.i32[name].assign((int32_t *)buf.data(), (int32_t *)buf.data() + count); break;
                case 1: out.f64[name].assign((double *)buf.data(), (double *)buf.data() + count); break;
                case 2: out.u8[name].assign(buf.data(), buf.data() + count); break;
                case 3: out.i8[name].assign((int8_t *)buf.data(), (int8_t *)buf.data() + count); break;
                default: std::fclose(f); throw std::runtime_error("ABMP: unknown dtype");
            }
        }
        std::fclose(f);
        return out;
    }

private:
    static void put(FILE *f, const std::string &name, uint32_t dtype, const void *data, size_t count, size_t esz) {
        char nm[24];
        std::memset(nm, 0, sizeof nm);
        std::strncpy(nm, name.c_str(), 23);
        uint64_t c = count;
        std::fwrite(nm, 1, 24, f);
        std::fwrite(&dtype, 4, 1, f);
        std::fwrite(&c, 8, 1, f);
        size_t bytes = count * esz, padded = (bytes + 7) & ~size_t(7);
        if (bytes) std::fwrite(data, 1, bytes, f);
        static const char zeros[8] = {0};
        if (padded > bytes) std::fwrite(zeros, 1, padded - bytes, f);
    }
};

}  // namespace abmcmc
