// Tiny tagged-table container shared by the oracle drivers (writer side).
// File = "SLRT" u32 count, then per entry: u32 nameLen, name, u32 dtype (0 = f32, 1 = u8), u64 n, payload.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

struct TableWriter {
    FILE* f;
    uint32_t count = 0;
    explicit TableWriter(const char* path) {
        f = fopen(path, "wb");
        if (!f) { perror(path); exit(1); }
        fwrite("SLRT", 1, 4, f);
        fwrite(&count, 4, 1, f);
    }
    void put(const std::string& name, uint32_t dtype, uint64_t n, const void* data) {
        uint32_t len = (uint32_t)name.size();
        fwrite(&len, 4, 1, f); fwrite(name.data(), 1, len, f);
        fwrite(&dtype, 4, 1, f); fwrite(&n, 8, 1, f);
        fwrite(data, dtype == 0 ? 4 : 1, n, f);
        ++count;
    }
    void floats(const std::string& name, const float* d, uint64_t n) { put(name, 0, n, d); }
    void bytes(const std::string& name, const void* d, uint64_t n) { put(name, 1, n, d); }
    ~TableWriter() { fseek(f, 4, SEEK_SET); fwrite(&count, 4, 1, f); fclose(f); }
};
