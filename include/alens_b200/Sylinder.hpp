// Sylinder.hpp -- rod record with the same fields, order and 568-byte layout as
// SimToolbox/Sylinder/Sylinder.hpp:38-84, so existing host code (protein binding, output writers) keeps
// reading/writing sy.pos, sy.orientation, sy.velCol ... unchanged and alens_set_rods_aos can take the
// container as is.
#ifndef ALENS_B200_SYLINDER_HPP_
#define ALENS_B200_SYLINDER_HPP_

#include <cmath>
#include <cstring>
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
