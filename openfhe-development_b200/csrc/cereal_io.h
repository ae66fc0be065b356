// cereal_io.h -- see cereal_io.cpp
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ofhe {
namespace cio {

struct DcrtInfo {
    uint32_t N, n_towers, format;  // format: 0 EVALUATION, 1 COEFFICIENT (the reference's enum Format)
    size_t header_end;             // offset just past the last tower and the polynomial's format word
};
// walks a serialized DCRTPoly; fills (optionally) moduli, roots (0 when a tower shares an earlier parameter object) and
// the byte offset of every tower's N raw words inside `buf`.  Throws std::runtime_error on malformed input.
DcrtInfo scan_dcrtpoly(const void* buf, size_t len, std::vector<uint64_t>* moduli, std::vector<uint64_t>* roots,
                       std::vector<size_t>* offsets);
// byte-identical to Serial::Serialize(DCRTPoly, SerType::BINARY) for a polynomial with these towers ([n_towers][N] words)
void write_dcrtpoly(uint32_t N, uint32_t n_towers, const uint64_t* moduli, const uint64_t* roots, uint32_t format,
                    const uint64_t* data, std::vector<uint8_t>& out);

}  // namespace cio
}  // namespace ofhe
