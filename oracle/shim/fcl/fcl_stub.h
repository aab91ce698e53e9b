/*
 * oracle/shim/fcl/fcl_stub.h - TEST INFRASTRUCTURE ONLY: stand-in for the FCL 0.7 types and calls the
 * reference's CollisionChecker uses (include/utils/collision_check.h:183-203,
 * src/utils/collision_check.cpp:17-246), so that the reference's own translation units compile in an
 * image without FCL (oracle/Makefile, `_ref`). FCL is a third-party, un-vendored dependency of the
 * reference; nothing is copied from it. `collide(shape, OcTree)` evaluates the occupied-voxel model of
 * ../../voxel_model.h - the same model the port uses - so a port-vs-_ref comparison pins the
 * reference's sampler control flow around the collision query, NOT the query itself (SURVEY 8a row
 * S4 stays "parity unpinned" beyond the reference's three FCL booleans). Distance queries
 * (CollisionChecker::getMinDistance, unused on the hot path) are not modelled.
 */
#pragma once
#include <Eigen/Dense>

#include <array>
#include <memory>
#include <vector>

#include "../octomap/octomap.h"
#include "../../voxel_model.h"

namespace fcl {

using Vector3f = Eigen::Vector3f;
using Matrix3f = Eigen::Matrix3f;
using Transform3f = Eigen::Isometry3f;

enum StubKind { KIND_BOX, KIND_CYLINDER, KIND_SPHERE, KIND_OCTREE };

template <typename S> class CollisionGeometry {
public:
  virtual ~CollisionGeometry() = default;
  virtual StubKind kind() const = 0;
  S cost_density = 1, threshold_occupied = 1;
};
using CollisionGeometryf = CollisionGeometry<float>;

template <typename S> class Box : public CollisionGeometry<S> {
public:
  S side[3];
  Box(S x, S y, S z) : side{x, y, z} {}
  StubKind kind() const override { return KIND_BOX; }
};
template <typename S> class Cylinder : public CollisionGeometry<S> {
public:
  S radius, lz;
  Cylinder(S r, S l) : radius(r), lz(l) {}
  StubKind kind() const override { return KIND_CYLINDER; }
};
template <typename S> class Sphere : public CollisionGeometry<S> {
public:
  S radius;
  explicit Sphere(S r) : radius(r) {}
  StubKind kind() const override { return KIND_SPHERE; }
};
using Boxf = Box<float>;
using Cylinderf = Cylinder<float>;
using Spheref = Sphere<float>;

template <typename S> class OcTree : public CollisionGeometry<S> {
public:
  std::shared_ptr<const octomap::OcTree> tree;
  explicit OcTree(const std::shared_ptr<const octomap::OcTree> &t) : tree(t) {}
  StubKind kind() const override { return KIND_OCTREE; }
  std::vector<std::array<S, 6>> toBoxes() const { return {}; }  // debug helper of the reference, unused
};
using OcTreef = OcTree<float>;

template <typename S> class CollisionObject {
public:
  std::shared_ptr<CollisionGeometry<S>> geom;
  Transform3f tf = Transform3f::Identity();
  explicit CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g) : geom(g) {}
  CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Transform3f &t) : geom(g), tf(t) {}
  CollisionObject(const std::shared_ptr<CollisionGeometry<S>> &g, const Matrix3f &R, const Vector3f &t) : geom(g) {
    tf.linear() = R;
    tf.translation() = t;
  }
  // (a derived-geometry shared_ptr converts implicitly, e.g. shared_ptr<OcTreef>)
  template <typename G, typename = std::enable_if_t<std::is_base_of<CollisionGeometry<S>, G>::value &&
                                                    !std::is_same<CollisionGeometry<S>, G>::value>>
  CollisionObject(const std::shared_ptr<G> &g, const Transform3f &t)
      : geom(std::static_pointer_cast<CollisionGeometry<S>>(g)), tf(t) {}
  void setTransform(const Transform3f &t) { tf = t; }
  void computeAABB() {}
  Vector3f getTranslation() const { return tf.translation(); }
  const Transform3f &getTransform() const { return tf; }
};
using CollisionObjectf = CollisionObject<float>;

template <typename S> struct CollisionResult {
  bool hit = false;
  bool isCollision() const { return hit; }
};
template <typename S> struct DefaultCollisionData {
  CollisionResult<S> result;
  bool done = false;
};
template <typename S> struct DistanceResult {
  S min_distance = 0;
};
template <typename S> struct DefaultDistanceData {
  DistanceResult<S> result;
  bool done = false;
};

namespace stub {
// shape (upright body at tf) against the occupied voxels of the octree object
inline bool shapeVsOctree(const CollisionObjectf &shape, const CollisionObjectf &oct) {
  const auto *tree = static_cast<const OcTreef *>(oct.geom.get());
  if (!tree->tree || tree->tree->points().empty()) return false;
  int vshape;
  float dims[3] = {0.0f, 0.0f, 0.0f};
  switch (shape.geom->kind()) {
  case KIND_CYLINDER: {
    const auto *c = static_cast<const Cylinderf *>(shape.geom.get());
    vshape = vox::VOX_CYLINDER;
    dims[0] = c->radius;
    dims[1] = c->lz;
    break;
  }
  case KIND_BOX: {
    const auto *b = static_cast<const Boxf *>(shape.geom.get());
    vshape = vox::VOX_BOX;
    dims[0] = b->side[0];
    dims[1] = b->side[1];
    dims[2] = b->side[2];
    break;
  }
  case KIND_SPHERE:
    vshape = vox::VOX_SPHERE;
    dims[0] = static_cast<const Spheref *>(shape.geom.get())->radius;
    break;
  default:
    return false;
  }
  orc::Iso3 stw;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) stw.L.m[i][j] = oct.tf.linear()(i, j);
    stw.t[i] = oct.tf.translation()(i);
  }
  // the world is rebuilt per query object: the reference builds a fresh octree per sensor update and
  // the cache below makes the many queries of one cycle share it
  struct Cache {
    const octomap::OcTree *tree = nullptr;
    uint64_t version = ~0ull;
    int shape = -1;
    float dims[3] = {0, 0, 0};
    orc::Iso3 stw{};
    vox::CollisionWorld W;
  };
  static thread_local Cache cache;
  bool same = cache.tree == tree->tree.get() && cache.version == tree->tree->version() && cache.shape == vshape;
  for (int i = 0; i < 3 && same; ++i) {
    same = cache.dims[i] == dims[i] && cache.stw.t[i] == stw.t[i];
    for (int j = 0; j < 3 && same; ++j) same = cache.stw.L.m[i][j] == stw.L.m[i][j];
  }
  if (!same) {
    cache.tree = tree->tree.get();
    cache.version = tree->tree->version();
    cache.shape = vshape;
    cache.stw = stw;
    for (int i = 0; i < 3; ++i) cache.dims[i] = dims[i];
    cache.W = vox::CollisionWorld();
    vox::initWorld(cache.W, vshape, dims, tree->tree->getResolution(), stw);
    if (vox::worldSupported(cache.W))
      for (const auto &p : tree->tree->points()) vox::insertPoint(cache.W, p.x(), p.y(), p.z());
  }
  if (!vox::worldSupported(cache.W)) return false;
  // upright body: position from the translation, heading from the rotation's first column
  const double x = shape.tf.translation()(0), y = shape.tf.translation()(1);
  const double yaw = std::atan2((double)shape.tf.linear()(1, 0), (double)shape.tf.linear()(0, 0));
  return vox::poseCollides(cache.W, x, y, yaw);
}
}  // namespace stub

template <typename S> bool DefaultCollisionFunction(CollisionObject<S> *o1, CollisionObject<S> *o2, void *data) {
  auto *cd = static_cast<DefaultCollisionData<S> *>(data);
  if (cd->done) return true;
  CollisionObject<S> *oct = o1->geom->kind() == KIND_OCTREE ? o1 : o2;
  CollisionObject<S> *shp = o1->geom->kind() == KIND_OCTREE ? o2 : o1;
  if (oct->geom->kind() == KIND_OCTREE && shp->geom->kind() != KIND_OCTREE && stub::shapeVsOctree(*shp, *oct)) {
    cd->result.hit = true;
    cd->done = true;
  }
  return cd->done;
}
template <typename S>
bool DefaultDistanceFunction(CollisionObject<S> *, CollisionObject<S> *, void *data, S &dist) {
  auto *dd = static_cast<DefaultDistanceData<S> *>(data);
  dd->result.min_distance = 0;  // not modelled (CollisionChecker::getMinDistance is off the hot path)
  dist = 0;
  return true;
}

template <typename S> class DynamicAABBTreeCollisionManager {
public:
  using CollisionCallBack = bool (*)(CollisionObject<S> *, CollisionObject<S> *, void *);
  using DistanceCallBack = bool (*)(CollisionObject<S> *, CollisionObject<S> *, void *, S &);
  void clear() { objs_.clear(); }
  void registerObject(CollisionObject<S> *o) { objs_.push_back(o); }
  void setup() {}
  void collide(CollisionObject<S> *obj, void *data, CollisionCallBack cb) const {
    for (auto *o : objs_)
      if (cb(o, obj, data)) return;
  }
  void distance(CollisionObject<S> *obj, void *data, DistanceCallBack cb) const {
    S d = 0;
    for (auto *o : objs_)
      if (cb(o, obj, data, d)) return;
  }

private:
  std::vector<CollisionObject<S> *> objs_;
};
using DynamicAABBTreeCollisionManagerf = DynamicAABBTreeCollisionManager<float>;

}  // namespace fcl
