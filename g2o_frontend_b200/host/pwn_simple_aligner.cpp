// pwn_simple_aligner -- the reference's CLI odometry driver (g2o_frontend/pwn_core/pwn_simple_aligner.cpp:28-269)
// re-expressed over the pwn:: classes of include/pwn/pwn.h, i.e. over the B200 library.
//
//   pwn_simple_aligner <config.conf> <out.jsonl> <depth0.pgm> <depth1.pgm> ...
//
// Reads a `key value` configuration (same keys as pwn_core/conf/pwn_aligner_1_1.conf), 16-bit binary PGM depth
// images in millimetres (the format of PlaneEx_gui/test_images/*.pgm), aligns every frame to the previous one and
// writes one JSON line per frame: the relative transform, the accumulated global transform (as t2v), inliers,
// error, correspondences and the image statistics of PwnMatcherBase::matchClouds.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

#include "pwn/pwn.h"
#include "pwn/boss_config.h"
#include "pwn/pyramid.h"
#include "pwn/tracker.h"

using namespace pwn;

static std::map<std::string, float> readConfig(const char *path) {  // pwn_simple_aligner.cpp:190-212
  std::map<std::string, float> m;
  std::ifstream is(path);
  std::string line;
  while (std::getline(is, line)) {
    if (line.empty() || line[0] == '#' || line[0] == '/') continue;
    std::istringstream ls(line);
    std::string key;
    float v;
    if (ls >> key >> v) m[key] = v;
  }
  return m;
}

static float get(const std::map<std::string, float> &m, const char *k, float def) {
  std::map<std::string, float>::const_iterator it = m.find(k);
  return it == m.end() ? def : it->second;
}

static bool readPgm16(const char *path, RawDepthImage &img) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  char magic[3] = {0, 0, 0};
  int w = 0, h = 0, maxv = 0;
  if (fscanf(f, "%2s", magic) != 1 || std::string(magic) != "P5") { fclose(f); return false; }
  int c = fgetc(f);
  while (c == '#' || c == '\n' || c == ' ' || c == '\r') {
    if (c == '#') while (c != '\n' && c != EOF) c = fgetc(f);
    c = fgetc(f);
  }
  ungetc(c, f);
  if (fscanf(f, "%d %d %d", &w, &h, &maxv) != 3) { fclose(f); return false; }
  fgetc(f);
  img.create(h, w);
  std::vector<unsigned char> buf((size_t)w * h * 2);
  size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size()) return false;
  for (size_t i = 0; i < (size_t)w * h; i++) img.buf[i] = (uint16_t)((buf[2 * i] << 8) | buf[2 * i + 1]);  // big-endian
  return true;
}

// --dump-config <file.conf>: parse a BOSS configuration of the reference's trackers (pwn_tracker2/conf/*.conf) into the
// pwn:: objects and print what they ended up with as one JSON object (no GPU is touched)
static int dumpBossConfig(const char *path) {
  PinholePointProjector alignerProjector, converterProjector;
  StatsCalculatorIntegralImage statsCalculator;
  PointInformationMatrixCalculator pointInfo;
  NormalInformationMatrixCalculator normalInfo;
  CorrespondenceFinder finder;
  Linearizer linearizer;
  Aligner aligner;
  Merger merger;
  VoxelCalculator voxel;
  BossPipeline p;
  p.alignerProjector = &alignerProjector; p.converterProjector = &converterProjector; p.statsCalculator = &statsCalculator;
  p.pointInformationMatrixCalculator = &pointInfo; p.normalInformationMatrixCalculator = &normalInfo;
  p.correspondenceFinder = &finder; p.linearizer = &linearizer; p.aligner = &aligner; p.merger = &merger;
  p.voxelCalculator = &voxel;
  std::vector<BossRecord> recs = bossLoad(path);
  configureFromBoss(recs, p);
  printf("{\"records\": %zu, \"outer_iterations\": %d, \"inner_iterations\": %d, \"inlier_max_chi2\": %.9g, \"robust_kernel\": %d, ",
         recs.size(), aligner.outerIterations(), aligner.innerIterations(), linearizer.inlierMaxChi2(), linearizer.robustKernel() ? 1 : 0);
  printf("\"inlier_distance_threshold\": %.9g, \"inlier_normal_angular_threshold\": %.9g, \"flat_curvature_threshold\": %.9g, "
         "\"inlier_curvature_ratio_threshold\": %.9g, ",
         finder.inlierDistanceThreshold(), finder.inlierNormalAngularThreshold(), finder.flatCurvatureThreshold(),
         finder.inlierCurvatureRatioThreshold());
  printf("\"K\": [");
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) printf("%s%.9g", r + c ? ", " : "", alignerProjector.cameraMatrix()(r, c));
  printf("], \"rows\": %d, \"cols\": %d, \"min_distance\": %.9g, \"max_distance\": %.9g, ", alignerProjector.imageRows(),
         alignerProjector.imageCols(), alignerProjector.minDistance(), alignerProjector.maxDistance());
  printf("\"reference_sensor_offset\": [");
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) printf("%s%.9g", r + c ? ", " : "", aligner.referenceSensorOffset().matrix()(r, c));
  printf("], \"world_radius\": %.9g, \"min_image_radius\": %d, \"max_image_radius\": %d, \"min_points\": %d, "
         "\"curvature_threshold\": %.9g, ",
         statsCalculator.worldRadius(), statsCalculator.minImageRadius(), statsCalculator.maxImageRadius(),
         statsCalculator.minPoints(), statsCalculator.curvatureThreshold());
  printf("\"flat_omega_p\": [%.9g, %.9g, %.9g], \"flat_omega_n\": [%.9g, %.9g, %.9g], ", pointInfo.flatInformationMatrix()(0, 0),
         pointInfo.flatInformationMatrix()(1, 1), pointInfo.flatInformationMatrix()(2, 2), normalInfo.flatInformationMatrix()(0, 0),
         normalInfo.flatInformationMatrix()(1, 1), normalInfo.flatInformationMatrix()(2, 2));
  printf("\"merger\": [%.9g, %.9g, %.9g], \"voxel_resolution\": %.9g, ", merger.distanceThreshold(), merger.normalThreshold(),
         merger.maxPointDepth(), voxel.resolution());
  printf("\"matcher_scale\": %d, \"frame_inlier_depth_threshold\": %.9g, \"new_frame_cloud_inliers_fraction\": %.9g, "
         "\"has_matcher\": %d, \"has_tracker\": %d}\n",
         p.matcherScale, p.frameInlierDepthThreshold, p.newFrameCloudInliersFraction, p.hasMatcher ? 1 : 0, p.hasTracker ? 1 : 0);
  return 0;
}

int main(int argc, char **argv) {
  if (argc == 3 && std::string(argv[1]) == "--dump-config") {
    try {
      return dumpBossConfig(argv[2]);
    } catch (const std::exception &e) {
      fprintf(stderr, "pwn_simple_aligner: %s\n", e.what());
      return 1;
    }
  }
  if (argc < 4) {
    fprintf(stderr, "usage: %s config.conf out.jsonl depth0.pgm depth1.pgm ...\n"
                    "       %s configurationFilename.txt depthImageListFilename.txt visualOdometryFilename.txt   (the reference's own)\n",
            argv[0], argv[0]);
    return 2;
  }
  // Exactly three arguments = the command line of the reference's pwn_simple_aligner (pwn_simple_aligner.cpp:33-40): a list of
  // "timestamp depthFilename" lines in, "timestamp x y z qx qy qz qw" lines out, and <depthFilename>.pwn next to every frame.
  const bool referenceCli = argc == 4;
  try {
    std::map<std::string, float> cfg = readConfig(argv[1]);
    // setInputParameters, pwn_simple_aligner.cpp:214-269
    int imageScale = (int)get(cfg, "imageScale", 1);
    const float depthScale = get(cfg, "depthScale", 0.001f);
    Matrix3f K;
    K.setIdentity();
    K(0, 0) = get(cfg, "fx", 525.0f); K(1, 1) = get(cfg, "fy", 525.0f);
    K(0, 2) = get(cfg, "cx", 319.5f); K(1, 2) = get(cfg, "cy", 239.5f);
    // pwn_simple_aligner.cpp:62-75 / pwn_aligner.cpp:74-87: tx ty tz qx qy qz qw are the INITIAL GLOBAL POSE of the
    // trajectory (a quaternion with its w, default (0,0,0,1), turned into a rotation matrix as it is, not normalised)
    Isometry3f initialT;
    {
      const float qx = get(cfg, "qx", 0), qy = get(cfg, "qy", 0), qz = get(cfg, "qz", 0), qw = get(cfg, "qw", 1.0f);
      // Eigen's QuaternionBase::toRotationMatrix
      const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
      const float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx;
      const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      Matrix3f R;
      R(0, 0) = 1.0f - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
      R(1, 0) = txy + twz; R(1, 1) = 1.0f - (txx + tzz); R(1, 2) = tyz - twx;
      R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1.0f - (txx + tyy);
      initialT.setLinear(R);
      initialT.setTranslation(get(cfg, "tx", 0), get(cfg, "ty", 0), get(cfg, "tz", 0));
    }
    // The reference's drivers run with an identity sensor offset (pwn_simple_aligner.cpp:126-127); the keys
    // sensorOffset{Tx,Ty,Tz,Qx,Qy,Qz} (an extension, v2t order) set one.
    Isometry3f sensorOffset;
    {
      Vector6f v;
      v(0) = get(cfg, "sensorOffsetTx", 0); v(1) = get(cfg, "sensorOffsetTy", 0); v(2) = get(cfg, "sensorOffsetTz", 0);
      v(3) = get(cfg, "sensorOffsetQx", 0); v(4) = get(cfg, "sensorOffsetQy", 0); v(5) = get(cfg, "sensorOffsetQz", 0);
      sensorOffset = v2t(v);
    }
    PinholePointProjector projector;
    projector.setMinDistance(get(cfg, "minDistance", 0.5f));
    projector.setMaxDistance(get(cfg, "maxDistance", 4.5f));
    projector.setCameraMatrix(K);

    StatsCalculatorIntegralImage statsCalculator;
    statsCalculator.setMinImageRadius((int)get(cfg, "minImageRadius", 10));
    statsCalculator.setMaxImageRadius((int)get(cfg, "maxImageRadius", 30));
    statsCalculator.setMinPoints((int)get(cfg, "minPoints", 50));
    statsCalculator.setCurvatureThreshold(get(cfg, "curvatureThreshold", 0.2f));
    statsCalculator.setWorldRadius(get(cfg, "worldRadius", 0.1f));
    PointInformationMatrixCalculator pointInformationMatrixCalculator;
    NormalInformationMatrixCalculator normalInformationMatrixCalculator;
    pointInformationMatrixCalculator.setCurvatureThreshold(get(cfg, "informationMatrixCurvatureThreshold", 0.02f));
    normalInformationMatrixCalculator.setCurvatureThreshold(get(cfg, "informationMatrixCurvatureThreshold", 0.02f));
    DepthImageConverterIntegralImage converter(&projector, &statsCalculator, &pointInformationMatrixCalculator,
                                               &normalInformationMatrixCalculator);

    CorrespondenceFinder correspondenceFinder;
    correspondenceFinder.setInlierDistanceThreshold(get(cfg, "inlierDistanceThreshold", 1.0f));
    correspondenceFinder.setInlierNormalAngularThreshold(get(cfg, "inlierNormalAngularThreshold", 0.95f));
    correspondenceFinder.setInlierCurvatureRatioThreshold(get(cfg, "inlierCurvatureRatioThreshold", 1.3f));
    correspondenceFinder.setFlatCurvatureThreshold(get(cfg, "flatCurvatureThreshold", 0.02f));
    Linearizer linearizer;
    linearizer.setInlierMaxChi2(get(cfg, "inlierMaxChi2", 9e3f));
    linearizer.setRobustKernel(get(cfg, "robustKernel", 1) != 0);
    Aligner aligner;
    aligner.setProjector(&projector);
    aligner.setLinearizer(&linearizer);
    aligner.setCorrespondenceFinder(&correspondenceFinder);
    aligner.setOuterIterations((int)get(cfg, "outerIterations", 10));
    aligner.setInnerIterations((int)get(cfg, "innerIterations", 1));
    aligner.setMinInliers((int)get(cfg, "minInliers", 100));
    aligner.setSensorOffset(sensorOffset);

    if (bossLooksLikeBoss(argv[1])) {
      // a BOSS file of the reference's trackers instead of a `key value` file: the records configure the same objects
      BossPipeline bp;
      bp.alignerProjector = &projector; bp.statsCalculator = &statsCalculator;
      bp.pointInformationMatrixCalculator = &pointInformationMatrixCalculator;
      bp.normalInformationMatrixCalculator = &normalInformationMatrixCalculator;
      bp.correspondenceFinder = &correspondenceFinder; bp.linearizer = &linearizer; bp.aligner = &aligner;
      configureFromBoss(bossLoad(argv[1]), bp);
      if (!bp.hasAligner) throw std::runtime_error("the BOSS file holds no Aligner record");
      // what the key/value file would have supplied comes from the records: camera matrix and image size of the aligner's
      // projector, the sensor offset of the aligner, the image scale of the matcher (PwnMatcherBase::scale)
      K = projector.cameraMatrix();
      sensorOffset = aligner.referenceSensorOffset();
      if (bp.hasMatcher) imageScale = bp.matcherScale;
      if (bp.hasTracker) cfg["newFrameInliersFraction"] = bp.newFrameCloudInliersFraction;
    }
    if (referenceCli) {
      // pwn_simple_aligner.cpp:117-187
      std::ifstream is(argv[2]);
      if (!is) throw std::runtime_error(std::string("Impossible to open depth image list file: ") + argv[2]);
      std::ofstream os(argv[3]);
      if (!os) throw std::runtime_error(std::string("Impossible to open visual odometry file: ") + argv[3]);
      converter.setKeepStats(true);  // the .pwn records carry the Stats
      // "timestamp x y z qx qy qz qw" with Quaternionf(globalT.linear()), normalize(): Eigen's matrix -> quaternion
      // (SURVEY.md Appendix A1), default ostream formatting like the reference
      auto writeOdometryLine = [&os](const std::string &timestamp, const Isometry3f &G) {
        const Matrix3f R = G.linear();
        float q[4];  // x y z w
        float t = (R(0, 0) + R(1, 1)) + R(2, 2);
        if (t > 0.0f) {
          t = sqrtf(t + 1.0f);
          q[3] = 0.5f * t;
          t = 0.5f / t;
          q[0] = (R(2, 1) - R(1, 2)) * t; q[1] = (R(0, 2) - R(2, 0)) * t; q[2] = (R(1, 0) - R(0, 1)) * t;
        } else {
          int i = 0;
          if (R(1, 1) > R(0, 0)) i = 1;
          if (R(2, 2) > R(i, i)) i = 2;
          const int j = (i + 1) % 3, k = (j + 1) % 3;
          t = sqrtf(R(i, i) - R(j, j) - R(k, k) + 1.0f);
          q[i] = 0.5f * t;
          t = 0.5f / t;
          q[3] = (R(k, j) - R(j, k)) * t; q[j] = (R(j, i) + R(i, j)) * t; q[k] = (R(k, i) + R(i, k)) * t;
        }
        const float qn = sqrtf(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        for (int i = 0; i < 4; i++) q[i] = q[i] / qn;
        os << timestamp << " " << G.translation().x() << " " << G.translation().y() << " " << G.translation().z() << " "
           << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
      };
      if (get(cfg, "localmap", 0) != 0) {
        // the reference's OTHER driver with the same command line, pwn_core/pwn_aligner.cpp:129-236: scene-based odometry
        // with the local map (the scene-NNN.pwn snapshots of :196-198,233-235 are not written: the device map does not
        // carry the per-point Stats those records hold)
        converter.setKeepGaussians(true);
        Merger merger;
        merger.setDepthImageConverter(&converter);
        merger.setDistanceThreshold(get(cfg, "mergerDistanceThreshold", 0.1f));
        merger.setNormalThreshold(get(cfg, "mergerNormalThreshold", cosf(10 * M_PI / 180.0f)));
        merger.setMaxPointDepth(get(cfg, "mergerMaxPointDepth", 10.0f));
        Cloud referenceScene, subscene;
        Isometry3f globalT = initialT, sceneT = initialT;
        const int chunkStep = (int)get(cfg, "chunkStep", 10);
        int counter = 0;
        bool firstDepth = true;
        while (is.good()) {
          char buf[1024];
          is.getline(buf, 1024);
          std::istringstream iss(buf);
          std::string timestamp, depthFilename;
          if (!(iss >> timestamp >> depthFilename)) continue;
          if (timestamp[0] == '#') continue;
          RawDepthImage raw;
          if (!readPgm16(depthFilename.c_str(), raw)) throw std::runtime_error("cannot read " + depthFilename);
          DepthImage scaledDepth;
          DepthImage_convertAndScale(scaledDepth, raw, imageScale, depthScale);
          if (firstDepth) {
            projector.setCameraMatrix(K);
            projector.setImageSize(raw.rows, raw.cols);
            projector.scale(1.0f / imageScale);
            correspondenceFinder.setImageSize(scaledDepth.rows, scaledDepth.cols);
            merger.setImageSize(scaledDepth.rows, scaledDepth.cols);
          }
          Cloud cloud;
          converter.compute(cloud, scaledDepth, sensorOffset);
          if (!firstDepth) {
            IntImage scaledIndexImage;
            projector.setTransform(sceneT * sensorOffset);
            projector.project(scaledIndexImage, scaledDepth, referenceScene);
            converter.setKeepGaussians(false);
            converter.compute(subscene, scaledDepth, sensorOffset);
            converter.setKeepGaussians(true);
            projector.setTransform(Isometry3f::Identity());
            aligner.setReferenceCloud(&subscene);
            aligner.setCurrentCloud(&cloud);
            aligner.setInitialGuess(Isometry3f::Identity());
            aligner.setSensorOffset(sensorOffset);
            aligner.align();
            globalT = globalT * aligner.T();
            globalT.fixLastRow();
            sceneT = sceneT * aligner.T();
            sceneT.fixLastRow();
          }
          if (!firstDepth && chunkStep > 0 && counter++ % chunkStep == 0) {
            sceneT = Isometry3f::Identity();
            referenceScene.clear();
          }
          referenceScene.add(cloud, sceneT);
          merger.merge(&referenceScene, sceneT * sensorOffset);
          projector.setTransform(Isometry3f::Identity());
          cloud.save((depthFilename + ".pwn").c_str(), globalT, 1, true);
          writeOdometryLine(timestamp, globalT);
          firstDepth = false;
        }
        return 0;
      }
      Cloud *cloud = 0, *previousCloud = 0;
      bool firstDepth = true;
      Isometry3f globalT = initialT;
      while (is.good()) {
        char buf[1024];
        is.getline(buf, 1024);
        std::istringstream iss(buf);
        std::string timestamp, depthFilename;
        if (!(iss >> timestamp >> depthFilename)) continue;
        if (timestamp[0] == '#') continue;
        RawDepthImage raw;
        if (!readPgm16(depthFilename.c_str(), raw)) throw std::runtime_error("cannot read " + depthFilename);
        DepthImage scaledDepth;
        DepthImage_convertAndScale(scaledDepth, raw, imageScale, depthScale);
        if (firstDepth) {
          projector.setCameraMatrix(K);
          projector.setImageSize(raw.rows, raw.cols);
          projector.scale(1.0f / imageScale);
          correspondenceFinder.setImageSize(scaledDepth.rows, scaledDepth.cols);
        }
        cloud = new Cloud();
        converter.compute(*cloud, scaledDepth, sensorOffset);
        if (!firstDepth) {
          aligner.setReferenceCloud(previousCloud);
          aligner.setCurrentCloud(cloud);
          aligner.setInitialGuess(Isometry3f::Identity());
          aligner.setSensorOffset(sensorOffset);
          aligner.align();
          globalT = globalT * aligner.T();
          globalT.fixLastRow();
          delete previousCloud;
        }
        cloud->save((depthFilename + ".pwn").c_str(), globalT, 1, true);
        writeOdometryLine(timestamp, globalT);
        previousCloud = cloud;
        firstDepth = false;
      }
      delete previousCloud;
      return 0;
    }
    FILE *out = fopen(argv[2], "w");
    if (!out) throw std::runtime_error("cannot open output file");
    if (get(cfg, "cloudio", 0) != 0) {
      // Cloud::save / load round trip (ASCII and binary .pwn) and Cloud::add on the first frame
      RawDepthImage raw;
      if (!readPgm16(argv[3], raw)) throw std::runtime_error("cannot read frame");
      DepthImage depth;
      DepthImage_convertAndScale(depth, raw, imageScale, depthScale);
      projector.setCameraMatrix(K);
      projector.setImageSize(raw.rows, raw.cols);
      projector.scale(1.0f / imageScale);
      converter.setKeepStats(true);
      Cloud cloud;
      converter.compute(cloud, depth, sensorOffset);
      std::string base(argv[2]);
      Isometry3f pose = sensorOffset, Ta, Tb;
      cloud.save((base + ".ascii.pwn").c_str(), pose, 1, false);
      cloud.save((base + ".bin.pwn").c_str(), pose, 1, true);
      Cloud a, b;
      bool okA = a.load(Ta, (base + ".ascii.pwn").c_str());
      bool okB = b.load(Tb, (base + ".bin.pwn").c_str());
      double maxAsciiErr = 0, maxBinErr = 0;
      for (size_t i = 0; i < cloud.size(); i++)
        for (int k = 0; k < 3; k++) {
          maxAsciiErr = std::max(maxAsciiErr, (double)std::fabs(a.points()[i][k] - cloud.points()[i][k]));
          maxBinErr = std::max(maxBinErr, (double)std::fabs(b.points()[i][k] - cloud.points()[i][k]));
          maxBinErr = std::max(maxBinErr, (double)std::fabs(b.normals()[i][k] - cloud.normals()[i][k]));
          maxBinErr = std::max(maxBinErr, (double)std::fabs(b.stats()[i](k, 3) - cloud.stats()[i](k, 3)));
        }
      size_t n0 = cloud.size();
      Isometry3f shift;
      shift.setTranslation(0.5f, 0.0f, 0.0f);
      Cloud sum;
      sum.add(cloud);
      sum.add(cloud, shift);
      double addErr = 0;
      for (size_t i = 0; i < n0; i++) {
        addErr = std::max(addErr, (double)std::fabs(sum.points()[n0 + i][0] - (cloud.points()[i][0] + 0.5f)));
        addErr = std::max(addErr, (double)std::fabs(sum.points()[i][1] - cloud.points()[i][1]));
        addErr = std::max(addErr, (double)std::fabs(sum.normals()[n0 + i][2] - cloud.normals()[i][2]));
      }
      fprintf(out, "{\"points\": %zu, \"ascii_ok\": %d, \"bin_ok\": %d, \"ascii_points\": %zu, \"bin_points\": %zu, "
                   "\"max_ascii_err\": %.9g, \"max_bin_err\": %.9g, \"sum_points\": %zu, \"add_err\": %.9g, "
                   "\"pose_err\": %.9g}\n",
              n0, okA ? 1 : 0, okB ? 1 : 0, a.size(), b.size(), maxAsciiErr, maxBinErr, sum.size(), addErr,
              (double)std::fabs(Tb.data()[12] - pose.data()[12]));
      fclose(out);
      return 0;
    }
    if (get(cfg, "stages", 0) != 0) {
      // the stage-level virtuals of the converter (StatsCalculator::compute, InformationMatrixCalculator::compute,
      // statscalculator.h:36, informationmatrixcalculator.h:83) chained by hand like
      // DepthImageConverterIntegralImage::compute does (depthimageconverterintegralimage.cpp:35-52), against the fused call
      RawDepthImage raw;
      if (!readPgm16(argv[3], raw)) throw std::runtime_error("cannot read frame");
      DepthImage depth;
      DepthImage_convertAndScale(depth, raw, imageScale, depthScale);
      projector.setCameraMatrix(K);
      projector.setImageSize(raw.rows, raw.cols);
      projector.scale(1.0f / imageScale);
      converter.setKeepStats(true);
      Cloud fused;
      converter.compute(fused, depth, Isometry3f::Identity());
      Cloud staged;
      IntImage indexImage;
      projector.setTransform(Isometry3f::Identity());
      projector.unProject(staged, indexImage, depth);
      projector.projectIntervals(statsCalculator.intervalImage(), depth, statsCalculator.worldRadius());
      NormalVector normals;
      StatsVector stats;
      statsCalculator.compute(normals, stats, ((const Cloud &)staged).points(), indexImage);
      InformationMatrixVector omegaP, omegaN;
      pointInformationMatrixCalculator.compute(omegaP, stats, normals);
      normalInformationMatrixCalculator.compute(omegaN, stats, normals);
      const Cloud &F = fused;
      double dN = 0, dC = 0, dP = 0, dNN = 0, dS = 0;
      size_t nonzero = 0;
      if (F.size() != normals.size()) throw std::runtime_error("stage-level point count differs");
      for (size_t i = 0; i < F.size(); i++) {
        for (int k = 0; k < 3; k++) dN = std::max(dN, (double)std::fabs(F.normals()[i][k] - normals[i][k]));
        if (normals[i][0] != 0 || normals[i][1] != 0 || normals[i][2] != 0) nonzero++;
        dC = std::max(dC, (double)std::fabs(F.stats()[i].curvature() - stats[i].curvature()));
        for (int k = 0; k < 16; k++) dS = std::max(dS, (double)std::fabs(F.stats()[i].m[k] - stats[i].m[k]));
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            dP = std::max(dP, (double)std::fabs(F.pointInformationMatrix()[i](r, c) - omegaP[i](r, c)));
            dNN = std::max(dNN, (double)std::fabs(F.normalInformationMatrix()[i](r, c) - omegaN[i](r, c)));
          }
      }
      fprintf(out, "{\"points\": %zu, \"nonzero_normals\": %zu, \"d_normals\": %.9g, \"d_curvature\": %.9g, \"d_stats\": %.9g, "
                   "\"d_omega_p\": %.9g, \"d_omega_n\": %.9g}\n", F.size(), nonzero, dN, dC, dS, dP, dNN);
      fclose(out);
      return 0;
    }
    if (get(cfg, "cloudcopy", 0) != 0) {
      // value semantics of pwn::Cloud and the host mirror: the same pair aligned (a) untouched, (b) after the current
      // cloud's host vectors were touched through a non-const accessor (host mirror -> re-upload, default keepStats = false),
      // (c) with by-value copies of both clouds (cloud.cpp:145, manifold_voronoi_extractor.cpp:82 copy clouds by value)
      if (argc < 5) throw std::runtime_error("cloudcopy needs two frames");
      Cloud clouds[2];
      for (int f = 0; f < 2; f++) {
        RawDepthImage raw;
        if (!readPgm16(argv[3 + f], raw)) throw std::runtime_error("cannot read frame");
        DepthImage depth;
        DepthImage_convertAndScale(depth, raw, imageScale, depthScale);
        if (f == 0) {
          projector.setCameraMatrix(K);
          projector.setImageSize(raw.rows, raw.cols);
          projector.scale(1.0f / imageScale);
          correspondenceFinder.setImageSize(depth.rows, depth.cols);
        }
        converter.compute(clouds[f], depth, sensorOffset);
      }
      Isometry3f Ts[3];
      int inl[3];
      double maxCurv = 0;
      for (int variant = 0; variant < 3; variant++) {
        Cloud copyRef, copyCur;
        const Cloud *ref = &clouds[0], *cur = &clouds[1];
        if (variant == 1) {
          Cloud &touched = clouds[1];
          const size_t n = touched.points().size();  // non-const accessor: the host mirror becomes the truth
          for (size_t i = 0; i < n; i++) maxCurv = std::max(maxCurv, (double)((const Cloud &)touched).stats()[i].curvature());
        } else if (variant == 2) {
          copyRef = clouds[0];        // host-valid source (copied vectors, uploaded on demand) or device-to-device
          Cloud byValue(clouds[1]);
          copyCur = byValue;
          ref = &copyRef;
          cur = &copyCur;
        }
        aligner.setReferenceCloud(const_cast<Cloud *>(ref));
        aligner.setCurrentCloud(const_cast<Cloud *>(cur));
        aligner.setInitialGuess(Isometry3f::Identity());
        aligner.setSensorOffset(sensorOffset);
        aligner.align();
        Ts[variant] = aligner.T();
        inl[variant] = aligner.inliers();
      }
      double d01 = 0, d02 = 0;
      for (int k = 0; k < 16; k++) {
        d01 = std::max(d01, (double)std::fabs(Ts[0].data()[k] - Ts[1].data()[k]));
        d02 = std::max(d02, (double)std::fabs(Ts[0].data()[k] - Ts[2].data()[k]));
      }
      fprintf(out, "{\"inliers\": [%d, %d, %d], \"dT_touched\": %.9g, \"dT_copied\": %.9g, \"max_mirror_curvature\": %.9g}\n",
              inl[0], inl[1], inl[2], d01, d02, maxCurv);
      fclose(out);
      return 0;
    }
    if (get(cfg, "localmap", 0) != 0) {
      // the scene-based odometry of pwn_core/pwn_aligner.cpp:140-215: every frame is aligned against the local map
      // re-rendered at the predicted pose, added to the map (Cloud::add) and fused into it (Merger::merge); the map is
      // voxelised at the end (VoxelCalculator)
      converter.setKeepGaussians(true);
      Merger merger;
      merger.setDepthImageConverter(&converter);
      merger.setDistanceThreshold(get(cfg, "mergerDistanceThreshold", 0.1f));
      merger.setNormalThreshold(get(cfg, "mergerNormalThreshold", cosf(10 * M_PI / 180.0f)));
      merger.setMaxPointDepth(get(cfg, "mergerMaxPointDepth", 10.0f));
      Cloud referenceScene, subscene;
      Isometry3f globalT = initialT, sceneT = initialT;  // pwn_aligner.cpp:137-140
      bool firstDepth = true;
      // pwn_aligner.cpp:72-73,194-205: after the first alignment and then every chunkStep frames the local map is closed
      // (the reference saves it as scene-%03d.pwn) and a new one is started from the current frame.  chunkStep 0 (where
      // the reference would divide by zero) = one map for the whole sequence.
      const int chunkStep = (int)get(cfg, "chunkStep", 10);
      int counter = 0;
      for (int a = 3; a < argc; a++) {
        RawDepthImage raw;
        if (!readPgm16(argv[a], raw)) throw std::runtime_error(std::string("cannot read ") + argv[a]);
        DepthImage scaledDepth;
        DepthImage_convertAndScale(scaledDepth, raw, imageScale, depthScale);
        if (firstDepth) {
          projector.setCameraMatrix(K);
          projector.setImageSize(raw.rows, raw.cols);
          projector.scale(1.0f / imageScale);
          correspondenceFinder.setImageSize(scaledDepth.rows, scaledDepth.cols);
          merger.setImageSize(scaledDepth.rows, scaledDepth.cols);
        }
        Cloud cloud;
        converter.compute(cloud, scaledDepth, sensorOffset);
        int inliers = 0;
        if (!firstDepth) {
          IntImage scaledIndexImage;
          projector.setTransform(sceneT * sensorOffset);
          projector.project(scaledIndexImage, scaledDepth, referenceScene);
          converter.setKeepGaussians(false);
          converter.compute(subscene, scaledDepth, sensorOffset);
          converter.setKeepGaussians(true);
          projector.setTransform(Isometry3f::Identity());
          aligner.setReferenceCloud(&subscene);
          aligner.setCurrentCloud(&cloud);
          aligner.setInitialGuess(Isometry3f::Identity());
          aligner.setSensorOffset(sensorOffset);
          aligner.align();
          inliers = aligner.inliers();
          globalT = globalT * aligner.T();
          sceneT = sceneT * aligner.T();
        }
        int newMap = 0;
        if (!firstDepth && chunkStep > 0 && counter++ % chunkStep == 0) {
          sceneT = Isometry3f::Identity();
          referenceScene.clear();
          newMap = 1;
        }
        const size_t before = referenceScene.size() + cloud.size();
        referenceScene.add(cloud, sceneT);
        merger.merge(&referenceScene, sceneT * sensorOffset);
        projector.setTransform(Isometry3f::Identity());
        fprintf(out, "{\"frame\": %d, \"inliers\": %d, \"added\": %zu, \"map_points\": %zu, \"new_map\": %d, \"globalT\": [", a - 3,
                inliers, before, referenceScene.size(), newMap);
        for (int i = 0; i < 16; i++) fprintf(out, "%s%.9g", i ? ", " : "", globalT.data()[i]);
        fprintf(out, "]}\n");
        firstDepth = false;
      }
      VoxelCalculator voxelCalculator;
      voxelCalculator.setResolution(get(cfg, "voxelResolution", 0.01f));
      const size_t n0 = referenceScene.size();
      voxelCalculator.compute(referenceScene);
      fprintf(out, "{\"map_points\": %zu, \"voxel_points\": %zu, \"has_gaussians\": %d}\n", n0, referenceScene.size(),
              referenceScene.hasGaussians() ? 1 : 0);
      fclose(out);
      return 0;
    }
    if (get(cfg, "tracker", 0) != 0) {
      // BASELINE config 3: PwnTracker::processFrame over the whole sequence (keyframe logic included)
      SequentialTracker tracker(&converter, &aligner);
      tracker.setScale(imageScale);
      tracker.setNewFrameInliersFraction(get(cfg, "newFrameInliersFraction", 0.4f));
      for (int a = 3; a < argc; a++) {
        RawDepthImage raw;
        if (!readPgm16(argv[a], raw)) throw std::runtime_error(std::string("cannot read ") + argv[a]);
        tracker.processFrame(raw, sensorOffset, K, Isometry3f::Identity(), depthScale);
        fprintf(out, "{\"frame\": %d, \"keyframe\": %d, \"keyframes\": %d, \"inliers\": %d, \"globalT\": [", a - 3,
                tracker.lastWasKeyframe() ? 1 : 0, tracker.numKeyframes(), tracker.lastInliers());
        for (int i = 0; i < 16; i++) fprintf(out, "%s%.9g", i ? ", " : "", tracker.globalT().data()[i]);
        fprintf(out, "]}\n");
      }
      fclose(out);
      return 0;
    }
    if (get(cfg, "pyramid", 0) != 0) {
      // BASELINE config 2: 3-level coarse-to-fine alignment of frame 1 against frame 0
      RawDepthImage r0, r1;
      if (!readPgm16(argv[3], r0) || !readPgm16(argv[4], r1)) throw std::runtime_error("cannot read the two frames");
      PyramidAligner pyr(&converter, &aligner);
      const int steps[3] = {4, 2, 1};
      for (int i = 0; i < 3; i++) {
        PyramidLevel L;
        L.step = steps[i];
        // pwn_aligner_1_4.conf radii at scale 4, pwn_aligner_1_1.conf at full resolution, in between at scale 2
        L.minImageRadius = steps[i] == 4 ? 3 : (steps[i] == 2 ? 5 : 10);
        L.maxImageRadius = steps[i] == 4 ? 6 : (steps[i] == 2 ? 15 : 30);
        L.minPoints = steps[i] == 4 ? 10 : (steps[i] == 2 ? 25 : 50);
        L.inlierDistanceThreshold = steps[i] == 1 ? 1.0f : 0.5f;
        L.outerIterations = (int)get(cfg, "outerIterations", 10);
        pyr.addLevel(L);
      }
      pyr.align(r0, r1, K, sensorOffset, Isometry3f::Identity(), depthScale);
      for (size_t li = 0; li < pyr.levelTransforms().size(); li++) {
        fprintf(out, "{\"level\": %zu, \"step\": %d, \"inliers\": %d, \"T\": [", li, steps[li], pyr.levelInliers()[li]);
        for (int i = 0; i < 16; i++) fprintf(out, "%s%.9g", i ? ", " : "", pyr.levelTransforms()[li].data()[i]);
        fprintf(out, "]}\n");
      }
      fclose(out);
      return 0;
    }
    Cloud *previous = 0;
    Isometry3f globalT = initialT;  // pwn_simple_aligner.cpp:128
    for (int a = 3; a < argc; a++) {
      RawDepthImage raw;
      if (!readPgm16(argv[a], raw)) throw std::runtime_error(std::string("cannot read ") + argv[a]);
      DepthImage depth;
      DepthImage_convertAndScale(depth, raw, imageScale, depthScale);
      // projector->scale(1/imageScale): pinholepointprojector.cpp:149-154
      projector.setCameraMatrix(K);
      projector.setImageSize(raw.rows, raw.cols);
      projector.scale(1.0f / imageScale);
      Cloud *current = new Cloud();
      converter.compute(*current, depth, sensorOffset);
      if (previous) {
        projector.setCameraMatrix(K);
        projector.setImageSize(raw.rows, raw.cols);
        projector.scale(1.0f / imageScale);
        correspondenceFinder.setImageSize(projector.imageRows(), projector.imageCols());
        aligner.setReferenceCloud(previous);
        aligner.setCurrentCloud(current);
        aligner.setInitialGuess(Isometry3f::Identity());
        aligner.align();
        globalT = globalT * aligner.T();
        Vector6f g = t2v(globalT);
        const nicp_align_result &r = aligner.lastResult();
        fprintf(out, "{\"frame\": %d, \"points\": %zu, \"T\": [", a - 3, current->size());
        for (int i = 0; i < 16; i++) fprintf(out, "%s%.9g", i ? ", " : "", aligner.T().data()[i]);
        fprintf(out, "], \"global\": [");
        for (int i = 0; i < 6; i++) fprintf(out, "%s%.9g", i ? ", " : "", g(i));
        fprintf(out, "], \"inliers\": %d, \"error\": %.9g, \"correspondences\": %d, \"image_nonZeros\": %d, "
                     "\"image_inliers\": %d, \"Hsum\": %.9g, \"time_ms\": %.3f}\n",
                aligner.inliers(), aligner.error(), correspondenceFinder.numCorrespondences(), r.image_non_zeros,
                r.image_inliers, linearizer.H()(0, 0) + linearizer.H()(5, 5), aligner.totalTime());
        delete previous;
      } else {
        fprintf(out, "{\"frame\": 0, \"points\": %zu}\n", current->size());
      }
      previous = current;
    }
    delete previous;
    fclose(out);
  } catch (const std::exception &e) {
    fprintf(stderr, "pwn_simple_aligner: %s\n", e.what());
    return 1;
  }
  return 0;
}
