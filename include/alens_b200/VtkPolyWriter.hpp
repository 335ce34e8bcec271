// VtkPolyWriter.hpp -- VTK XML PolyData (.vtp / .pvtp) output in the wire format the reference produces, so that its
// post-processing (ParaView states, Examples/*/*.pvsm, the Verify.py scripts) reads our results unchanged:
//   * one line cell per record (two points), arrays as base64 "binary" DataArrays with a UInt32 byte count in front,
//     header and payload encoded separately (SimToolbox/Util/IOHelper.hpp:231-251, Util/Base64.hpp:248-262)
//   * piece files `<prefix><Name>_r<rank>_<postfix>.vtp` + one parallel index `<prefix><Name>_<postfix>.pvtp`
// Table driven: a file is a list of named, typed columns; Sylinder::writeVTP and ConstraintCollector::writeVTP only fill
// the columns (reference field lists: Sylinder/Sylinder.hpp:185-453, Constraint/ConstraintCollector.cpp:76-224).
// Host-side output code: not on the hot path.
#ifndef ALENS_B200_VTKPOLYWRITER_HPP_
#define ALENS_B200_VTKPOLYWRITER_HPP_

#include <cstdint>
#include <fstream>
#include <string>
#include <type_traits>
#include <vector>

namespace alens_vtk {

inline void base64Append(const unsigned char *in, size_t n, std::string &out) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    size_t i = 0;
    out.reserve(out.size() + 4 * ((n + 2) / 3));
    for (; i + 2 < n; i += 3) {
        const unsigned v = (in[i] << 16) | (in[i + 1] << 8) | in[i + 2];
        out.push_back(T[(v >> 18) & 63]); out.push_back(T[(v >> 12) & 63]);
        out.push_back(T[(v >> 6) & 63]);  out.push_back(T[v & 63]);
    }
    if (n - i == 2) {
        const unsigned v = (in[i] << 16) | (in[i + 1] << 8);
        out.push_back(T[(v >> 18) & 63]); out.push_back(T[(v >> 12) & 63]); out.push_back(T[(v >> 6) & 63]); out.push_back('=');
    } else if (n - i == 1) {
        const unsigned v = in[i] << 16;
        out.push_back(T[(v >> 18) & 63]); out.push_back(T[(v >> 12) & 63]); out.push_back('='); out.push_back('=');
    }
}

template <class T>
const char *typeName() {
    return std::is_same<T, int32_t>::value ? "Int32" : std::is_same<T, float>::value ? "Float32"
         : std::is_same<T, double>::value ? "Float64" : "UInt8";
}

/// one named column: type, components per tuple, payload already encoded (UInt32 byte count, then the data)
struct Column {
    std::string name, type, b64;
    int ncomp = 1;
    template <class T>
    static Column of(const std::string &name, int ncomp, const std::vector<T> &v) {
        static_assert(std::is_same<T, int32_t>::value || std::is_same<T, float>::value || std::is_same<T, double>::value ||
                          std::is_same<T, uint8_t>::value, "VTK column type");
        Column c;
        c.name = name;
        c.type = typeName<T>();
        c.ncomp = ncomp;
        const uint32_t bytes = (uint32_t)(v.size() * sizeof(T));
        base64Append(reinterpret_cast<const unsigned char *>(&bytes), 4, c.b64);
        base64Append(reinterpret_cast<const unsigned char *>(v.data()), bytes, c.b64);
        return c;
    }
    void write(std::ofstream &f) const {
        f << "<DataArray Name=\"" << name << "\" type=\"" << type << "\" NumberOfComponents=\"" << ncomp
          << "\" format=\"binary\">\n" << b64 << "\n</DataArray>\n";
    }
};

/// a .vtp piece whose cells are nLines two-point lines over 2 nLines points (ends = 6 doubles per line)
inline void writeLinePiece(const std::string &path, int nLines, const std::vector<double> &ends,
                           const std::vector<Column> &pointData, const std::vector<Column> &cellData) {
    std::vector<int32_t> conn(2 * (size_t)nLines), offs((size_t)nLines);
    for (int i = 0; i < nLines; i++) {
        conn[2 * i] = 2 * i;
        conn[2 * i + 1] = 2 * i + 1;
        offs[i] = 2 * i + 2;
    }
    std::ofstream f(path, std::ios::out);
    f << "<?xml version=\"1.0\"?>\n"
      << "<VTKFile type=\"PolyData\" version=\"1.0\" byte_order=\"LittleEndian\"  header_type=\"UInt32\">\n"
      << "<PolyData>\n";
    f << "<Piece NumberOfPoints=\"" << nLines * 2 << "\" NumberOfLines=\"" << nLines << "\">\n";
    f << "<Points>\n";
    Column::of("position", 3, ends).write(f);
    f << "</Points>\n<Lines>\n";
    Column::of("connectivity", 1, conn).write(f);
    Column::of("offsets", 1, offs).write(f);
    f << "</Lines>\n<PointData Scalars=\"scalars\">\n";
    for (const auto &c : pointData) c.write(f);
    f << "</PointData>\n<CellData Scalars=\"scalars\">\n";
    for (const auto &c : cellData) c.write(f);
    f << "</CellData>\n</Piece>\n</PolyData>\n</VTKFile>" << std::endl;
}

/// the .pvtp index: declares the columns (name, type, components) and lists the piece files
struct Field {
    std::string name, type;
    int ncomp;
};
inline void writeParallelIndex(const std::string &path, const std::vector<Field> &pointFields,
                               const std::vector<Field> &cellFields, const std::vector<std::string> &pieces) {
    std::ofstream f(path, std::ios::out);
    f << "<?xml version=\"1.0\"?>\n"
      << "<VTKFile type=\"PPolyData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt32\"> \n"
      << "<PPolyData GhostLevel=\"0\"> \n";
    auto decl = [&](const char *tag, const std::vector<Field> &fs) {
        f << "<" << tag << " Scalars=\"scalars\">\n";
        for (const auto &d : fs)
            f << "<PDataArray Name=\"" << d.name << "\" type=\"" << d.type << "\" NumberOfComponents=\"" << d.ncomp
              << "\" format=\"binary\"/>\n";
        f << "</" << tag << ">\n";
    };
    decl("PPointData", pointFields);
    decl("PCellData", cellFields);
    f << "<PPoints> \n<PDataArray NumberOfComponents=\"3\" type=\"Float64\" format=\"binary\"/>\n</PPoints> \n";
    for (const auto &p : pieces) f << "<Piece Source=\"" << p << "\"/>\n";
    f << "</PPolyData>\n</VTKFile>\n";
}

} // namespace alens_vtk
#endif
