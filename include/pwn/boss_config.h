// BOSS configuration files of the reference's trackers, read into the pwn:: classes of pwn/pwn.h.
//
// g2o_frontend/pwn_boss serialises every pwn:: object as one record `"ClassName" { ...json... }` carrying an integer
// "#id"; pointers between objects are `{ "#pointer" : id }` (pwn_boss/aligner.cpp:12-47, pinholepointprojector.cpp:10-24,
// depthimageconverter.cpp:20-49, ...; the files live in pwn_tracker2/conf/).  Eigen matrices are `{ "values" : [...] }`
// in ROW-major order (boss_map/eigen_boss_plugin.hpp:1-33), poses are t2v 6-vectors.  This header is a small
// self-contained reader (no BOSS, no JSON library): bossLoad() parses the records, configureFromBoss() applies the
// first Aligner / DepthImageConverter(IntegralImage) / Merger / VoxelCalculator of a file to already constructed
// objects through their setters -- what pwn_boss's deserialize() methods do.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pwn.h"

namespace pwn {

struct BossValue {
  enum Type { Null, Number, Bool, String, Array, Object } type;
  double number;
  bool boolean;
  std::string string;
  std::vector<BossValue> array;
  std::vector<std::pair<std::string, BossValue> > object;
  BossValue() : type(Null), number(0), boolean(false) {}

  const BossValue *find(const std::string &key) const {
    for (size_t i = 0; i < object.size(); i++)
      if (object[i].first == key) return &object[i].second;
    return 0;
  }
  const BossValue &at(const std::string &key) const {
    const BossValue *v = find(key);
    if (!v) throw std::runtime_error("BOSS: missing field \"" + key + "\"");
    return *v;
  }
  double num(const std::string &key) const {
    const BossValue &v = at(key);
    if (v.type == Bool) return v.boolean ? 1.0 : 0.0;
    if (v.type != Number) throw std::runtime_error("BOSS: field \"" + key + "\" is not a number");
    return v.number;
  }
  double num(const std::string &key, double def) const { return find(key) ? num(key) : def; }
  // { "#pointer" : id } -> id, -1 if absent / null
  int pointer(const std::string &key) const {
    const BossValue *v = find(key);
    if (!v || v->type != Object) return -1;
    const BossValue *p = v->find("#pointer");
    return p && p->type == Number ? (int)p->number : -1;
  }
  // { "values" : [...] }
  std::vector<float> values(const std::string &key) const {
    const BossValue &a = at(key).at("values");
    std::vector<float> out;
    for (size_t i = 0; i < a.array.size(); i++) out.push_back((float)a.array[i].number);
    return out;
  }
};

struct BossRecord {
  std::string className;
  int id;
  BossValue fields;
};

namespace boss_detail {
struct Parser {
  const std::string &s;
  size_t i;
  explicit Parser(const std::string &text) : s(text), i(0) {}
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
  bool end() { ws(); return i >= s.size(); }
  void fail(const char *what) const {
    std::ostringstream os;
    os << "BOSS parse error at offset " << i << ": " << what;
    throw std::runtime_error(os.str());
  }
  std::string str() {
    if (s[i] != '"') fail("expected a string");
    std::string out;
    for (i++; i < s.size() && s[i] != '"'; i++) {
      if (s[i] == '\\' && i + 1 < s.size()) {
        char c = s[++i];
        out += c == 'n' ? '\n' : c == 't' ? '\t' : c;
      } else {
        out += s[i];
      }
    }
    if (i >= s.size()) fail("unterminated string");
    i++;
    return out;
  }
  BossValue value() {
    ws();
    if (i >= s.size()) fail("unexpected end of file");
    BossValue v;
    char c = s[i];
    if (c == '{') {
      v.type = BossValue::Object;
      i++;
      ws();
      if (i < s.size() && s[i] == '}') { i++; return v; }
      for (;;) {
        ws();
        std::string key = str();
        ws();
        if (i >= s.size() || s[i] != ':') fail("expected ':'");
        i++;
        v.object.push_back(std::make_pair(key, value()));
        ws();
        if (i < s.size() && s[i] == ',') { i++; continue; }
        if (i < s.size() && s[i] == '}') { i++; break; }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      v.type = BossValue::Array;
      i++;
      ws();
      if (i < s.size() && s[i] == ']') { i++; return v; }
      for (;;) {
        v.array.push_back(value());
        ws();
        if (i < s.size() && s[i] == ',') { i++; continue; }
        if (i < s.size() && s[i] == ']') { i++; break; }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.type = BossValue::String;
      v.string = str();
    } else if (s.compare(i, 4, "true") == 0) {
      v.type = BossValue::Bool; v.boolean = true; i += 4;
    } else if (s.compare(i, 5, "false") == 0) {
      v.type = BossValue::Bool; v.boolean = false; i += 5;
    } else if (s.compare(i, 4, "null") == 0) {
      i += 4;
    } else {
      // strtod is as lenient as BOSS's own reader (hand-edited values such as `000` occur in the reference's files)
      const char *b = s.c_str() + i;
      char *e = 0;
      v.type = BossValue::Number;
      v.number = std::strtod(b, &e);
      if (e == b) fail("expected a value");
      i += (size_t)(e - b);
    }
    return v;
  }
};
}  // namespace boss_detail

inline std::vector<BossRecord> bossParse(const std::string &text) {
  std::vector<BossRecord> out;
  boss_detail::Parser p(text);
  while (!p.end()) {
    BossRecord r;
    r.className = p.str();
    r.fields = p.value();
    if (r.fields.type != BossValue::Object) p.fail("record body must be an object");
    const BossValue *id = r.fields.find("#id");
    r.id = id && id->type == BossValue::Number ? (int)id->number : -1;
    out.push_back(r);
  }
  return out;
}
inline std::vector<BossRecord> bossLoad(const char *path) {
  std::ifstream is(path);
  if (!is) throw std::runtime_error(std::string("cannot open ") + path);
  std::stringstream ss;
  ss << is.rdbuf();
  return bossParse(ss.str());
}
// is this file in the BOSS record format (as opposed to the `key value` files of pwn_core/conf)?
inline bool bossLooksLikeBoss(const char *path) {
  std::ifstream is(path);
  char c;
  while (is.get(c))
    if (!std::isspace((unsigned char)c)) return c == '"';
  return false;
}

namespace boss_detail {
inline const BossRecord *firstOf(const std::vector<BossRecord> &recs, const char *cls) {
  for (size_t i = 0; i < recs.size(); i++)
    if (recs[i].className == cls) return &recs[i];
  return 0;
}
inline const BossRecord *byId(const std::vector<BossRecord> &recs, int id) {
  for (size_t i = 0; i < recs.size(); i++)
    if (recs[i].id == id && id >= 0) return &recs[i];
  return 0;
}
inline Isometry3f pose(const BossValue &f, const char *key) {
  std::vector<float> v = f.values(key);
  if (v.size() != 6) throw std::runtime_error(std::string("BOSS: ") + key + " must hold 6 values");
  Vector6f t;
  for (int k = 0; k < 6; k++) t(k) = v[k];
  return v2t(t);
}
inline InformationMatrix info(const BossValue &f, const char *key) {
  std::vector<float> v = f.values(key);
  if (v.size() != 16) throw std::runtime_error(std::string("BOSS: ") + key + " must hold 16 values");
  InformationMatrix m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = v[4 * r + c];  // row-major (eigen_boss_plugin.hpp)
  return m;
}
}  // namespace boss_detail

// pwn_boss/pointprojector.cpp:23-35 + pinholepointprojector.cpp:18-24
inline void bossConfigure(PinholePointProjector &p, const BossValue &f) {
  p.setTransform(boss_detail::pose(f, "transform"));
  p.setMinDistance((float)f.num("minDistance"));
  p.setMaxDistance((float)f.num("maxDistance"));
  p.setImageSize((int)f.num("imageRows"), (int)f.num("imageCols"));
  std::vector<float> k = f.values("cameraMatrix");
  if (k.size() != 9) throw std::runtime_error("BOSS: cameraMatrix must hold 9 values");
  Matrix3f K;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) K(r, c) = k[3 * r + c];
  p.setCameraMatrix(K);
  p.setBaseline((float)f.num("baseline", 0.075));
  p.setAlpha((float)f.num("alpha", 0.1));
}
// pwn_boss/statscalculatorintegralimage.cpp:19-27
inline void bossConfigure(StatsCalculatorIntegralImage &s, const BossValue &f) {
  s.setWorldRadius((float)f.num("worldRadius"));
  s.setMaxImageRadius((int)f.num("imageMaxRadius"));
  s.setMinImageRadius((int)f.num("imageMinRadius"));
  s.setMinPoints((int)f.num("minPoints"));
  s.setCurvatureThreshold((float)f.num("curvatureThreshold"));
}
// pwn_boss/informationmatrixcalculator.cpp:15-19
inline void bossConfigure(InformationMatrixCalculator &c, const BossValue &f) {
  c.setFlatInformationMatrix(boss_detail::info(f, "flatInformationMatrix"));
  c.setNonFlatInformationMatrix(boss_detail::info(f, "nonflatInformationMatrix"));
}
// pwn_boss/correspondencefinder.cpp:20-28
inline void bossConfigure(CorrespondenceFinder &c, const BossValue &f) {
  c.setInlierDistanceThreshold((float)f.num("inlierDistanceThreshold"));
  c.setFlatCurvatureThreshold((float)f.num("flatCurvatureThreshold"));
  c.setInlierCurvatureRatioThreshold((float)f.num("inlierCurvatureRatioThreshold"));
  c.setInlierNormalAngularThreshold((float)f.num("inlierNormalAngularThreshold"));
  c.setImageSize((int)f.num("rows"), (int)f.num("cols"));
}
// pwn_boss/linearizer.cpp:22-26
inline void bossConfigure(Linearizer &l, const BossValue &f) {
  l.setRobustKernel(f.num("robustKernel") != 0.0);
  l.setInlierMaxChi2((float)f.num("inlierMaxChi2"));
}
// pwn_boss/merger.cpp:17-21, voxelcalculator.cpp:14-16
inline void bossConfigure(Merger &m, const BossValue &f) {
  m.setDistanceThreshold((float)f.num("distanceThreshold"));
  m.setNormalThreshold((float)f.num("normalThreshold"));
  m.setMaxPointDepth((float)f.num("maxPointDepth"));
}
inline void bossConfigure(VoxelCalculator &v, const BossValue &f) { v.setResolution((float)f.num("resolution")); }

// The first Aligner of the file with the Linearizer / CorrespondenceFinder / projector it points to, and the first
// DepthImageConverter(IntegralImage) with its projector / statistics / information-matrix calculators, applied to the
// caller's objects (pwn_boss/aligner.cpp:36-47, depthimageconverter.cpp:43-49).  Absent records leave the object as it is.
struct BossPipeline {
  PinholePointProjector *alignerProjector, *converterProjector;
  StatsCalculatorIntegralImage *statsCalculator;
  PointInformationMatrixCalculator *pointInformationMatrixCalculator;
  NormalInformationMatrixCalculator *normalInformationMatrixCalculator;
  CorrespondenceFinder *correspondenceFinder;
  Linearizer *linearizer;
  Aligner *aligner;
  Merger *merger;
  VoxelCalculator *voxelCalculator;
  // outputs: PwnMatcherBase (pwn_tracker2/pwn_matcher_base.cpp:20-36) and PwnTracker records, if present
  bool hasAligner, hasMatcher, hasTracker;
  int matcherScale;
  float frameInlierDepthThreshold, newFrameCloudInliersFraction;
  int minCloudInliers, frameMinNonZeroThreshold, frameMaxOutliersThreshold, frameMinInliersThreshold;
  BossPipeline()
      : alignerProjector(0), converterProjector(0), statsCalculator(0), pointInformationMatrixCalculator(0),
        normalInformationMatrixCalculator(0), correspondenceFinder(0), linearizer(0), aligner(0), merger(0), voxelCalculator(0),
        hasAligner(false), hasMatcher(false), hasTracker(false), matcherScale(1), frameInlierDepthThreshold(50.0f),
        newFrameCloudInliersFraction(0.4f), minCloudInliers(0), frameMinNonZeroThreshold(0), frameMaxOutliersThreshold(0),
        frameMinInliersThreshold(0) {}
};
inline void configureFromBoss(const std::vector<BossRecord> &recs, BossPipeline &p) {
  using namespace boss_detail;
  if (const BossRecord *mt = firstOf(recs, "PwnMatcherBase")) {
    p.hasMatcher = true;
    p.matcherScale = (int)mt->fields.num("scale");
    p.frameInlierDepthThreshold = (float)mt->fields.num("frameInlierDepthThreshold");
  }
  if (const BossRecord *tr = firstOf(recs, "PwnTracker")) {
    p.hasTracker = true;
    p.newFrameCloudInliersFraction = (float)tr->fields.num("newFrameCloudInliersFraction", 0.4);
    p.minCloudInliers = (int)tr->fields.num("minCloudInliers", 0);
    p.frameMinNonZeroThreshold = (int)tr->fields.num("frameMinNonZeroThreshold", 0);
    p.frameMaxOutliersThreshold = (int)tr->fields.num("frameMaxOutliersThreshold", 0);
    p.frameMinInliersThreshold = (int)tr->fields.num("frameMinInliersThreshold", 0);
  }
  if (const BossRecord *al = firstOf(recs, "Aligner")) {
    p.hasAligner = true;
    if (p.aligner) {
      p.aligner->setOuterIterations((int)al->fields.num("outerIterations"));
      p.aligner->setInnerIterations((int)al->fields.num("innerIterations"));
      p.aligner->setReferenceSensorOffset(pose(al->fields, "referenceSensorOffset"));
      p.aligner->setCurrentSensorOffset(pose(al->fields, "currentSensorOffset"));
    }
    const BossRecord *r;
    if (p.alignerProjector && (r = byId(recs, al->fields.pointer("projector"))) && r->className == "PinholePointProjector")
      bossConfigure(*p.alignerProjector, r->fields);
    if (p.linearizer && (r = byId(recs, al->fields.pointer("linearizer")))) bossConfigure(*p.linearizer, r->fields);
    if (p.correspondenceFinder && (r = byId(recs, al->fields.pointer("correspondenceFinder"))))
      bossConfigure(*p.correspondenceFinder, r->fields);
  }
  const BossRecord *cv = firstOf(recs, "DepthImageConverterIntegralImage");
  if (!cv) cv = firstOf(recs, "DepthImageConverter");
  if (cv) {
    const BossRecord *r;
    if (p.converterProjector && (r = byId(recs, cv->fields.pointer("pointProjector"))) && r->className == "PinholePointProjector")
      bossConfigure(*p.converterProjector, r->fields);
    if (p.statsCalculator && (r = byId(recs, cv->fields.pointer("statsCalculator"))) &&
        r->className == "StatsCalculatorIntegralImage")
      bossConfigure(*p.statsCalculator, r->fields);
    if (p.pointInformationMatrixCalculator && (r = byId(recs, cv->fields.pointer("pointInfoCalculator"))))
      bossConfigure(*p.pointInformationMatrixCalculator, r->fields);
    if (p.normalInformationMatrixCalculator && (r = byId(recs, cv->fields.pointer("normalInfoCalculator"))))
      bossConfigure(*p.normalInformationMatrixCalculator, r->fields);
  }
  if (const BossRecord *m = firstOf(recs, "Merger"))
    if (p.merger) bossConfigure(*p.merger, m->fields);
  if (const BossRecord *v = firstOf(recs, "VoxelCalculator"))
    if (p.voxelCalculator) bossConfigure(*p.voxelCalculator, v->fields);
}

}  // namespace pwn
