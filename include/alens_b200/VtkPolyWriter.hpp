// VtkPolyWriter.hpp -- VTK XML PolyData (.vtp / .pvtp) output (and, for restarts, input) in the wire format the reference produces, so that its
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
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
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

// ------------------------------------------------------------------------------------------------------------------
// Reader for the same wire format: what SylinderSystem::setInitialFromVTKFile (SylinderSystem.cpp:406-476) takes from
// vtkXMLPPolyDataReader when a run is restarted -- the points and the cell arrays of every piece of a .pvtp index,
// concatenated in piece order.  Only the encoding the writers above (and the reference's IOHelper) produce is read:
// format="binary" (base64, UInt32 byte count encoded in front of the payload), uncompressed, little endian.

inline std::vector<unsigned char> base64Decode(const std::string &in, size_t begin, size_t end) {
    static const std::vector<int> T = [] {
        std::vector<int> t(256, -1);
        const char *a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) t[(unsigned char)a[i]] = i;
        return t;
    }();
    std::vector<unsigned char> out;
    out.reserve((end - begin) / 4 * 3);
    unsigned acc = 0;
    int bits = 0;
    for (size_t i = begin; i < end; i++) {
        const int v = T[(unsigned char)in[i]];
        if (v < 0) continue; // '=' padding and white space
        acc = (acc << 6) | (unsigned)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((unsigned char)((acc >> bits) & 0xff));
        }
    }
    return out;
}

/// one DataArray, widened to double (Int32 / UInt8 / Float32 / Float64)
struct Array {
    int ncomp = 1;
    std::vector<double> v;
    size_t tuples() const { return ncomp > 0 ? v.size() / (size_t)ncomp : 0; }
};
struct PolyData {
    std::vector<double> points; ///< 3 per point
    std::vector<std::pair<std::string, Array>> pointData, cellData;
    const Array &cell(const std::string &name) const {
        for (const auto &c : cellData)
            if (c.first == name) return c.second;
        throw std::runtime_error("vtk: no cell array named " + name);
    }
    size_t numberOfPoints() const { return points.size() / 3; }
};

inline std::string xmlAttr(const std::string &tag, const std::string &key) {
    const size_t a = tag.find(key + "=\"");
    if (a == std::string::npos) return "";
    const size_t b = a + key.size() + 2, e = tag.find('"', b);
    return e == std::string::npos ? "" : tag.substr(b, e - b);
}

/// appends the arrays of one .vtp piece to `out` (arrays are matched by name, in file order for the first piece)
inline void readPiece(const std::string &path, PolyData &out) {
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f) throw std::runtime_error("vtk: cannot open " + path);
    const std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto sectionOf = [&](size_t at) { // which of <Points> / <PointData> / <CellData> / other encloses position `at`
        const char *names[3] = {"<Points>", "<PointData", "<CellData"};
        const char *ends[3] = {"</Points>", "</PointData>", "</CellData>"};
        for (int k = 0; k < 3; k++) {
            const size_t a = s.rfind(names[k], at);
            if (a == std::string::npos) continue;
            const size_t e = s.find(ends[k], a);
            if (e != std::string::npos && e > at) return k;
        }
        return 3;
    };
    size_t at = 0;
    while ((at = s.find("<DataArray", at)) != std::string::npos) {
        const size_t close = s.find('>', at);
        if (close == std::string::npos) break;
        const std::string tag = s.substr(at, close - at + 1);
        const size_t endTag = s.find("</DataArray>", close);
        if (endTag == std::string::npos) throw std::runtime_error("vtk: unterminated DataArray in " + path);
        const int sec = sectionOf(at);
        at = endTag;
        if (sec == 3) continue; // connectivity / offsets
        if (xmlAttr(tag, "format") != "binary") throw std::runtime_error("vtk: only format=\"binary\" is read: " + path);
        const std::string type = xmlAttr(tag, "type"), name = xmlAttr(tag, "Name");
        const std::string nc = xmlAttr(tag, "NumberOfComponents");
        // the UInt32 byte count is encoded on its own: 4 bytes -> 8 characters
        size_t b = close + 1;
        while (b < endTag && (s[b] == '\n' || s[b] == '\r' || s[b] == ' ' || s[b] == '\t')) b++;
        if (endTag - b < 8) throw std::runtime_error("vtk: short DataArray " + name);
        const std::vector<unsigned char> head = base64Decode(s, b, b + 8);
        if (head.size() < 4) throw std::runtime_error("vtk: bad DataArray header " + name);
        const uint32_t bytes = (uint32_t)head[0] | ((uint32_t)head[1] << 8) | ((uint32_t)head[2] << 16) | ((uint32_t)head[3] << 24);
        const std::vector<unsigned char> raw = base64Decode(s, b + 8, endTag);
        if (raw.size() < bytes) throw std::runtime_error("vtk: truncated DataArray " + name);
        Array a;
        a.ncomp = nc.empty() ? 1 : std::atoi(nc.c_str());
        auto widen = [&](auto zero) {
            using T = decltype(zero);
            const size_t n = bytes / sizeof(T);
            a.v.resize(n);
            for (size_t i = 0; i < n; i++) {
                T t;
                std::memcpy(&t, raw.data() + i * sizeof(T), sizeof(T));
                a.v[i] = (double)t;
            }
        };
        if (type == "Float64") widen(double(0));
        else if (type == "Float32") widen(float(0));
        else if (type == "Int32") widen(int32_t(0));
        else if (type == "UInt8") widen(uint8_t(0));
        else throw std::runtime_error("vtk: unsupported DataArray type " + type);
        if (sec == 0) {
            out.points.insert(out.points.end(), a.v.begin(), a.v.end());
            continue;
        }
        auto &list = sec == 1 ? out.pointData : out.cellData;
        bool found = false;
        for (auto &c : list)
            if (c.first == name) {
                c.second.v.insert(c.second.v.end(), a.v.begin(), a.v.end());
                found = true;
            }
        if (!found) list.emplace_back(name, std::move(a));
    }
}

/// a .pvtp index: every <Piece Source="..."/> (relative to the index), merged in order; a .vtp is read by itself
inline PolyData readParallel(const std::string &path) {
    PolyData out;
    if (path.size() > 4 && path.compare(path.size() - 4, 4, ".vtp") == 0) {
        readPiece(path, out);
        return out;
    }
    std::ifstream f(path);
    if (!f) throw std::runtime_error("vtk: cannot open " + path);
    const std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const size_t slash = path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
    size_t at = 0;
    while ((at = s.find("<Piece", at)) != std::string::npos) {
        const size_t close = s.find('>', at);
        if (close == std::string::npos) break;
        const std::string src = xmlAttr(s.substr(at, close - at + 1), "Source");
        if (!src.empty()) readPiece(dir + src, out);
        at = close;
    }
    return out;
}

} // namespace alens_vtk
#endif
