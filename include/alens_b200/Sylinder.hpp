// Sylinder.hpp -- rod record with the same fields, order and 568-byte layout as
// SimToolbox/Sylinder/Sylinder.hpp:38-84, so existing host code (protein binding, output writers) keeps
// reading/writing sy.pos, sy.orientation, sy.velCol ... unchanged and alens_set_rods_aos can take the
// container as is.
#ifndef ALENS_B200_SYLINDER_HPP_
#define ALENS_B200_SYLINDER_HPP_

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "VtkPolyWriter.hpp"
#include <cstdio>
#include <limits>

#ifndef GEO_INVALID_INDEX
#define GEO_INVALID_INDEX (-1)
#endif

struct Link {
    int prev = GEO_INVALID_INDEX, next = GEO_INVALID_INDEX;
};

class Sylinder {
  public:
    int gid = GEO_INVALID_INDEX;
    int globalIndex = GEO_INVALID_INDEX;
    int rank = -1;
    int group = -1;
    bool isImmovable = false;
    double radius = 0, radiusCollision = 0, length = 0, lengthCollision = 0, radiusSearch = 0, sepmin = 0, colBuf = 0;
    double pos[3] = {0, 0, 0};
    double orientation[4] = {0, 0, 0, 1}; ///< quaternion (x,y,z,w); direction = orientation * (0,0,1)
    double vel[3], omega[3], velCol[3], omegaCol[3], velBi[3], omegaBi[3], velNonB[3], omegaNonB[3];
    double force[3], torque[3], forceCol[3], torqueCol[3], forceBi[3], torqueBi[3], forceNonB[3], torqueNonB[3];
    double velBrown[3], omegaBrown[3];

    Sylinder() { clear(); }
    Sylinder(int gid_, double radius_, double radiusCollision_, double length_, double lengthCollision_,
             const double pos_[3] = nullptr, const double orientation_[4] = nullptr)
        : gid(gid_), radius(radius_), radiusCollision(radiusCollision_), length(length_),
          lengthCollision(lengthCollision_) {
        if (pos_) std::memcpy(pos, pos_, sizeof(pos));
        if (orientation_) std::memcpy(orientation, orientation_, sizeof(orientation));
        clear();
    }
    void clear() { // Sylinder.cpp:34-59
        std::memset(vel, 0, (char *)(omegaBrown + 3) - (char *)vel);
        sepmin = std::numeric_limits<double>::max();
        globalIndex = GEO_INVALID_INDEX;
        rank = -1;
    }
    bool isSphere(bool collision = false) const {
        return collision ? lengthCollision < radiusCollision * 2 : length < radius * 2;
    }
    /// direction = orientation * (0,0,1) (Eigen quaternion rotation, SylinderNear.hpp:86)
    void direction(double d[3]) const {
        const double x = orientation[0], y = orientation[1], z = orientation[2], w = orientation[3];
        const double ux = y + y, uy = -(x + x);
        d[0] = w * ux + (-(z * uy));
        d[1] = w * uy + z * ux;
        d[2] = 1.0 + (x * uy - y * ux);
    }
    /// Sylinder::stepEuler (Sylinder.cpp:91-99) with EquatnHelper::rotateEquatn (Util/EquatnHelper.hpp:74-90) on the host:
    /// used when a run is restarted (the device copy is stepped by alens_step_euler)
    void stepEuler(double dt) {
        for (int k = 0; k < 3; k++) pos[k] += vel[k] * dt;
        const double w = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
        if (w < std::numeric_limits<float>::epsilon()) return;
        const double winv = 1 / w, sw = std::sin(w * dt / 2), cw = std::cos(w * dt / 2);
        const double s = orientation[3], p[3] = {orientation[0], orientation[1], orientation[2]};
        const double cr[3] = {omega[1] * p[2] - omega[2] * p[1], omega[2] * p[0] - omega[0] * p[2], omega[0] * p[1] - omega[1] * p[0]};
        double q[4];
        for (int k = 0; k < 3; k++) q[k] = s * sw * omega[k] * winv + cw * p[k] + sw * winv * cr[k];
        q[3] = s * cw - (p[0] * omega[0] + p[1] * omega[1] + p[2] * omega[2]) * sw * winv;
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int k = 0; k < 4; k++) orientation[k] = q[k] / n;
    }
    /// Eigen's quaternion * vector (QuaternionBase::_transformVector): uv = q.vec x v; uv += uv; v + w uv + q.vec x uv
    void rotate(const double v[3], double out[3]) const {
        const double x = orientation[0], y = orientation[1], z = orientation[2], w = orientation[3];
        double ux = y * v[2] - z * v[1], uy = z * v[0] - x * v[2], uz = x * v[1] - y * v[0];
        ux += ux; uy += uy; uz += uz;
        out[0] = v[0] + w * ux + (y * uz - z * uy);
        out[1] = v[1] + w * uy + (z * ux - x * uz);
        out[2] = v[2] + w * uz + (x * uy - y * ux);
    }

    /// the columns of Sylinder_r<rank>_<postfix>.vtp in the reference's order (Sylinder.hpp:185-259, :281-453)
    struct VtkColumn {
        const char *name;
        int ncomp;
        const char *type;
    };
    static const std::vector<VtkColumn> &vtkCellColumns() {
        static const std::vector<VtkColumn> cols = {
            {"gid", 1, "Int32"}, {"group", 1, "Int32"}, {"isImmovable", 1, "UInt8"}, {"radius", 1, "Float32"},
            {"radiusCollision", 1, "Float32"}, {"length", 1, "Float32"}, {"lengthCollision", 1, "Float32"},
            {"vel", 3, "Float32"}, {"omega", 3, "Float32"}, {"velCollision", 3, "Float32"}, {"omegaCollision", 3, "Float32"},
            {"velBilateral", 3, "Float32"}, {"omegaBilateral", 3, "Float32"}, {"velNonBrown", 3, "Float32"},
            {"omegaNonBrown", 3, "Float32"}, {"force", 3, "Float32"}, {"torque", 3, "Float32"},
            {"forceCollision", 3, "Float32"}, {"torqueCollision", 3, "Float32"}, {"forceBilateral", 3, "Float32"},
            {"torqueBilateral", 3, "Float32"}, {"forceNonBrown", 3, "Float32"}, {"torqueNonBrown", 3, "Float32"},
            {"velBrown", 3, "Float32"}, {"omegaBrown", 3, "Float32"}, {"xnorm", 3, "Float32"}, {"znorm", 3, "Float32"}};
        return cols;
    }
    /// Sylinder::writeVTP: one line per rod between its two end points, `prefix` ends in '/' (a folder)
    template <class Container>
    static void writeVTP(const Container &sylinder, const int sylinderNumber, const std::string &prefix,
                         const std::string &postfix, int rank) {
        using alens_vtk::Column;
        const size_t n = (size_t)sylinderNumber;
        std::vector<double> ends(6 * n);
        std::vector<uint8_t> label(2 * n), imm(n);
        std::vector<int32_t> gid(n), group(n);
        std::vector<float> scal[4];
        for (auto &v : scal) v.resize(n);
        // the 3-vectors of the record, in file order: member pointers into the 568-byte layout
        double (Sylinder::*const vecs[])[3] = {&Sylinder::vel, &Sylinder::omega, &Sylinder::velCol, &Sylinder::omegaCol,
                                               &Sylinder::velBi, &Sylinder::omegaBi, &Sylinder::velNonB, &Sylinder::omegaNonB,
                                               &Sylinder::force, &Sylinder::torque, &Sylinder::forceCol, &Sylinder::torqueCol,
                                               &Sylinder::forceBi, &Sylinder::torqueBi, &Sylinder::forceNonB,
                                               &Sylinder::torqueNonB, &Sylinder::velBrown, &Sylinder::omegaBrown};
        constexpr int NV = sizeof(vecs) / sizeof(vecs[0]);
        std::vector<float> v3[NV + 2];
        for (auto &v : v3) v.resize(3 * n);
        for (size_t i = 0; i < n; i++) {
            const Sylinder &sy = sylinder[i];
            const double ex[3] = {1, 0, 0}, ez[3] = {0, 0, 1};
            double nx[3], nz[3];
            sy.rotate(ex, nx);
            sy.rotate(ez, nz);
            for (int k = 0; k < 3; k++) {
                ends[6 * i + k] = sy.pos[k] - nz[k] * (sy.length * 0.5);
                ends[6 * i + 3 + k] = sy.pos[k] + nz[k] * (sy.length * 0.5);
                for (int a = 0; a < NV; a++) v3[a][3 * i + k] = (float)(sy.*vecs[a])[k];
                v3[NV][3 * i + k] = (float)nx[k];
                v3[NV + 1][3 * i + k] = (float)nz[k];
            }
            label[2 * i] = 0;
            label[2 * i + 1] = 1;
            gid[i] = sy.gid;
            group[i] = sy.group;
            imm[i] = sy.isImmovable ? 1 : 0;
            scal[0][i] = (float)sy.radius; scal[1][i] = (float)sy.radiusCollision;
            scal[2][i] = (float)sy.length; scal[3][i] = (float)sy.lengthCollision;
        }
        const auto &cols = vtkCellColumns();
        std::vector<Column> cell;
        cell.push_back(Column::of(cols[0].name, 1, gid));
        cell.push_back(Column::of(cols[1].name, 1, group));
        cell.push_back(Column::of(cols[2].name, 1, imm));
        for (int a = 0; a < 4; a++) cell.push_back(Column::of(cols[3 + a].name, 1, scal[a]));
        for (int a = 0; a < NV + 2; a++) cell.push_back(Column::of(cols[7 + a].name, 3, v3[a]));
        alens_vtk::writeLinePiece(prefix + "Sylinder_r" + std::to_string(rank) + "_" + postfix + ".vtp", sylinderNumber, ends,
                                  {Column::of("endLabel", 1, label)}, cell);
    }
    /// Sylinder::writePVTP: the parallel index over nProcs piece files
    static void writePVTP(const std::string &prefix, const std::string &postfix, const int nProcs) {
        std::vector<alens_vtk::Field> cell;
        for (const auto &c : vtkCellColumns()) cell.push_back({c.name, c.type, c.ncomp});
        std::vector<std::string> pieces;
        for (int i = 0; i < nProcs; i++) pieces.push_back("Sylinder_r" + std::to_string(i) + "_" + postfix + ".vtp");
        alens_vtk::writeParallelIndex(prefix + "Sylinder_" + postfix + ".pvtp", {{"endLabel", "UInt8", 1}}, cell, pieces);
    }

    /// one line of SylinderAscii_*.dat, the format the reference reads back (Sylinder.cpp:101-109,
    /// SylinderSystem.cpp:317-344): `C|S gid radius minus[3] plus[3] group`
    void writeAscii(FILE *fptr) const {
        double d[3];
        direction(d);
        const char typeChar = isImmovable ? 'S' : 'C';
        std::fprintf(fptr, "%c %d %.8g %.8g %.8g %.8g %.8g %.8g %.8g %d\n", typeChar, gid, radius, //
                     pos[0] - 0.5 * length * d[0], pos[1] - 0.5 * length * d[1], pos[2] - 0.5 * length * d[2],
                     pos[0] + 0.5 * length * d[0], pos[1] + 0.5 * length * d[1], pos[2] + 0.5 * length * d[2], group);
    }
};

/// header of SylinderAscii_*.dat (Sylinder.hpp:459-464)
class SylinderAsciiHeader {
  public:
    int nparticle = 0;
    double time = 0;
    void writeAscii(FILE *fp) const { std::fprintf(fp, "%d \n %lf\n", nparticle, time); }
};

static_assert(sizeof(Sylinder) == 568, "Sylinder record must keep the reference layout");

#endif
