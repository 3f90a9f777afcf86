// Host entry points of the mocap ingest + device forward kinematics (drt_skeleton.cu).
#pragma once
#include <cstddef>
#include <string>

namespace drt {
struct Skeleton;
// Parses ASF + AMC text, uploads, runs skeleton_fk over every frame on `device`.  Returns a drt_status.
int skeletonCreate(const char* asf, size_t asf_len, const char* amc, size_t amc_len, double scale, int device, Skeleton** out);
void skeletonDestroy(Skeleton* s);
int skeletonCylinders(const Skeleton* s);          // bones without the root = cylinders per frame
int skeletonFrames(const Skeleton* s);
float skeletonFkMs(const Skeleton* s);             // device time of the forward-kinematics kernel (CUDA events)
const double* skeletonHostTable(const Skeleton* s);   // [frame][cylinder][6], copied back from the device table once
int skeletonReadBones(const Skeleton* s, int frame0, int n, double* out);   // device table -> host
// parse only (host): bone count incl. the root, frame count, parent index and DOF bit mask (rx ry rz tx ty tz = bits 0..5) per bone
int skeletonParseInfo(const char* asf, size_t asf_len, const char* amc, size_t amc_len, double scale, int* n_bones, int* n_frames,
                      int* parents, int* dofs, int cap);
const std::string& skeletonError();
}  // namespace drt
