// oracle/stubs/zoltan_dd_cpp.h -- TEST INFRASTRUCTURE ONLY: one-process stand-in for Zoltan's distributed directory
// (Zoltan_DD) as SimToolbox/Trilinos/ZDD.hpp uses it: gid -> fixed-size user data, Update then Find.
#pragma once
#include "zoltan_types.h"
#include <cstring>
#include <mpi.h>
#include <unordered_map>
#include <vector>
class Zoltan_DD {
  public:
    int Create(MPI_Comm, int, int, int userLen, int, int) {
        len_ = userLen;
        return ZOLTAN_OK;
    }
    int Update(ZOLTAN_ID_PTR gid, ZOLTAN_ID_PTR, char *data, int *, int count) {
        for (int i = 0; i < count; i++) tbl_[gid[i]].assign(data + (size_t)i * len_, data + (size_t)(i + 1) * len_);
        return ZOLTAN_OK;
    }
    int Find(ZOLTAN_ID_PTR gid, ZOLTAN_ID_PTR, char *data, int *, int count, int *owner) {
        int rc = ZOLTAN_OK;
        for (int i = 0; i < count; i++) {
            auto it = tbl_.find(gid[i]);
            if (it == tbl_.end()) { rc = ZOLTAN_WARN; continue; }
            memcpy(data + (size_t)i * len_, it->second.data(), len_);
            if (owner) owner[i] = 0;
        }
        return rc;
    }
    void Print() const {}
    void Stats() const {}

  private:
    int len_ = 0;
    std::unordered_map<ZOLTAN_ID_TYPE, std::vector<char>> tbl_;
};
