/*
 * oracle/shim/octomap/octomap.h - TEST INFRASTRUCTURE ONLY: stand-in for the few octomap types the
 * reference's CollisionChecker touches (include/utils/collision_check.h:91-136,183-203), so that the
 * reference's own translation units compile in an image without octomap (oracle/Makefile, `_ref`).
 * octomap is a third-party, un-vendored dependency of the reference; nothing is copied from it. The
 * tree only records the inserted end points: the occupied-voxel semantics live in ../../voxel_model.h.
 */
#pragma once
#include <cstdint>
#include <vector>

namespace octomap {

struct point3d {
  float v[3];
  point3d(float x = 0.0f, float y = 0.0f, float z = 0.0f) : v{x, y, z} {}
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
};

class Pointcloud {
public:
  std::vector<point3d> pts;
  void push_back(float x, float y, float z) { pts.emplace_back(x, y, z); }
  void push_back(const point3d &p) { pts.push_back(p); }
  void clear() { pts.clear(); }
  size_t size() const { return pts.size(); }
};

class OcTree {
public:
  explicit OcTree(double resolution) : res_(resolution) {}
  void clear() { pts_.clear(); ++version_; }
  void setResolution(double r) { res_ = r; ++version_; }
  double getResolution() const { return res_; }
  // after a clear() the occupied leaves are the voxels of the end points (see voxel_model.h)
  void insertPointCloud(const Pointcloud &cloud, const point3d & /*sensor_origin*/) {
    pts_.insert(pts_.end(), cloud.pts.begin(), cloud.pts.end());
    ++version_;
  }
  const std::vector<point3d> &points() const { return pts_; }
  uint64_t version() const { return version_; }

private:
  double res_;
  std::vector<point3d> pts_;
  uint64_t version_ = 0;
};

}  // namespace octomap
