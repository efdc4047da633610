// batch_multi_bev_gen — drop-in replacement of the reference tool of the same name
// (soytony/Point-Cloud-Preprocessing-Tools, BatchMultiBevGen.cpp:664-771): same positional arguments
// `[keyframes_root_dir] [sensor_type]`, same inputs (keyframe_point_cloud/*.pcd, keyframe_pose.csv), same outputs
// (non_ground_point_cloud/, output_multi_bev/{binary,image}/, output_single_bev/{csv,image}/, keyframe_label.csv) and
// the same progress lines on stdout.  The per-frame arithmetic runs on B200 GPUs through the C-ABI of
// libbevgen_cuda.so (include/bevgen.h); this file is only host plumbing:
//   loader pool (PCD parse)  ->  per-GPU worker (pack into pinned SoA, bevgen_process_host)  ->  encode pool
//   (bin / 25 PNG / CSV / PCD), so file encoding is off the GPU critical path.  Frames shard over GPUs by batch index;
//   label rows are split per GPU and gathered on the host; no collective.
// Extra trailing options (not in the reference): --gpus N, --workers-per-gpu W, --batch B, --threads T, --no-encode, --no-pcd,
//   --png-level L, --json-metrics FILE, --no-packed (parse every PCD on the host instead of de-interleaving binary
//   payloads on the GPU).
// Compiled a second time with -DBATCH_CLOUD_MANIP as `batch_cloud_manip <keyframes_root_dir>` (BatchCloudManip.cpp:269-331,
// SURVEY 8(f)-3): same ordering + ground removal with the HDL-64E shape hard-coded there (:12-13, :85), the 201x201 float
// bird-view map of saveAsMat (:201-239) into output_bvm/<name>.csv/.png, and non_ground_point_cloud/<name>.pcd; no labels.
#include <dirent.h>
#include <malloc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <filesystem>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <iterator>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bevgen.h"
#include "image_io.h"
#include "pcd_io.h"

namespace fs = std::filesystem;
using Clock = std::chrono::steady_clock;

// ---------------------------------------------------------------------------------------------------------------
struct ThreadPool {
  std::vector<std::thread> th; std::deque<std::function<void()>> q; std::mutex mu; std::condition_variable cv, idle_cv;
  bool stop = false; int active = 0;
  explicit ThreadPool(int n) {
    for (int i = 0; i < n; i++) th.emplace_back([this] {
      for (;;) {
        std::function<void()> f;
        { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return stop || !q.empty(); }); if (q.empty()) return; f = std::move(q.front()); q.pop_front(); active++; }
        f();
        { std::lock_guard<std::mutex> l(mu); active--; if (q.empty() && active == 0) idle_cv.notify_all(); }
      }
    });
  }
  void submit(std::function<void()> f) { { std::lock_guard<std::mutex> l(mu); q.push_back(std::move(f)); } cv.notify_one(); }
  void submit_urgent(std::function<void()> f) { { std::lock_guard<std::mutex> l(mu); q.push_front(std::move(f)); } cv.notify_one(); }
  // fn(0) .. fn(n-1) on the pool, ahead of everything queued (a GPU worker is waiting for them), the caller taking its share
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    struct St { std::atomic<int> next{0}, done{0}; std::mutex m; std::condition_variable c; };
    auto st = std::make_shared<St>();
    auto body = [st, n, &fn] {
      for (;;) { const int i = st->next++; if (i >= n) break; fn(i); if (++st->done == n) { std::lock_guard<std::mutex> l(st->m); st->c.notify_all(); } }
    };
    const int helpers = std::min<int>(n - 1, (int)th.size());
    for (int h = 0; h < helpers; h++) submit_urgent([st, n, &fn] {      // a helper that arrives late finds nothing left and never touches fn
      for (;;) { const int i = st->next++; if (i >= n) break; fn(i); if (++st->done == n) { std::lock_guard<std::mutex> l(st->m); st->c.notify_all(); } }
    });
    body();
    std::unique_lock<std::mutex> l(st->m);
    st->c.wait(l, [&] { return st->done.load() == n; });
  }
  void wait_idle() { std::unique_lock<std::mutex> l(mu); idle_cv.wait(l, [&] { return q.empty() && active == 0; }); }
  ~ThreadPool() { { std::lock_guard<std::mutex> l(mu); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
};

struct Options {
  std::string root, sensor;
  int gpus = 1, batch = 16, threads = 0, png_level = 1;
  int workers_per_gpu = 1;   // host threads (each with its own context) feeding one GPU: a worker stages, calls the GPU and hands
                             // its batch to the encode pool one after the other, so a second one overlaps those serial sections
  bool encode = true, write_pcd = true, packed = true;
  bool bvm_mode = false;     // batch_cloud_manip
  std::string json_metrics;
};

struct Dirs { std::string pcd_in, non_ground, multi_bin, multi_img, single_csv, single_img, pose_file, label_file, bvm; };

[[maybe_unused]] static void usage_and_exit(const char* argv0) {   // BatchMultiBevGen.cpp:666-689
  std::cout << "Usage: " << argv0 << " [keyframes_root_dir] [sensor_type]\n\n"
            << "[keyframes_root_dir] should be organized as follows: \n"
            << "[keyframes_root_dir]\n"
            << "├ keyframe_point_cloud/ <- folder for selected point clouds in pcd format for each frame \n"
            << "├ keyframe_pose.csv <- 6-DoF pose for each frame \n"
            << "└ keyframe_pose_format.csv <- 6-DoF pose format description \n"
            << "\n"
            << "[sensor_type] could be HDL_32E, HDL_64E or OS1_64. \n"
            << "\n"
            << "This binary generates ground-removed point clouds, single & multi layer BEV images and creates geometric "
               "distance-based labels for each point cloud. After running the binary, you will have files organized as follows: \n "
            << "[keyframes_root_dir]\n "
            << "├ ... \n "
            << "├ non_ground_point_cloud/ <- folder for ground-removes point clouds in pcd format \n "
            << "├ output_multi_bev/ <- folder for multi-layer BEV images \n "
            << "└ output_single_bev <- folder for single-layer BEV images \n "
            << std::endl;
  exit(1);
}

static void reset_dir(const std::string& d) {   // `rm -rf` + `mkdir -p` (BatchMultiBevGen.cpp:49-70, :704-705) without forking a shell
  std::error_code ec;
  fs::remove_all(d, ec);
  fs::create_directories(d, ec);
}

// getPcdFileNames, BatchMultiBevGen.cpp:469-494: suffix after the last '.' must be "pcd"; lexicographic sort.
static std::vector<std::string> list_pcd(const std::string& path) {
  std::vector<std::string> out;
  DIR* d = opendir(path.c_str());
  if (!d) { std::cerr << "Folder doesn't Exist!" << std::endl; return out; }
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name;
    if (n == "." || n == "..") continue;
    if (n.substr(n.find_last_of('.') + 1) != "pcd") continue;
    out.push_back(path.back() == '/' ? path + n : path + "/" + n);
  }
  closedir(d);
  std::sort(out.begin(), out.end());
  return out;
}

static std::string short_name_of(const std::string& f) {   // :739-742
  int start_pos = (int)f.find_last_of('/') + 1;
  int end_pos = (int)f.find_last_of('.') - 1;
  return f.substr(start_pos, end_pos - start_pos + 1);
}

// readKeyframePose, BatchMultiBevGen.cpp:381-460: whitespace-separated entries, each split on ','; exactly 16 tokens
// or stop; x,y,z = float(std::stod(token)).  Only x,y,z feed the label stage.
static std::vector<float> read_poses(const std::string& file) {
  std::ifstream f(file);
  if (f.is_open()) std::cout << "loaded keyframe pose file: " << file << std::endl;
  else { std::cerr << "failed to load keyframe pose file: " << file << std::endl; exit(1); }
  std::vector<float> xyz;
  std::string entry;
  while (f >> entry) {
    std::vector<std::string> tok; std::stringstream ss(entry); std::string s;
    while (getline(ss, s, ',')) tok.push_back(s);
    if (tok.size() != 16) { std::cerr << "Size of entry_token is: " << tok.size() << ", while expecting 16. " << std::endl; break; }
    try {
      for (int k = 1; k <= 3; k++) xyz.push_back((float)std::stod(tok[k]));
      for (int k = 7; k < 16; k++) (void)std::stod(tok[k]);     // the reference parses the rotation too (and would throw here)
    } catch (const std::exception& e) { std::cerr << "bad number in keyframe pose file: " << e.what() << std::endl; exit(1); }
  }
  std::cout << "Finish reading all keyframe pose, total " << xyz.size() / 3 << " entries. " << std::endl;
  return xyz;
}

// ---------------------------------------------------------------------------------------------------------------
struct PinnedSet {     // pinned staging of one batch: SoA inputs + outputs
  size_t cap_pts = 0; int cap_frames = 0; size_t S = 0;
  float *x = 0, *y = 0, *z = 0, *inten = 0; uint16_t *row = 0, *col = 0; int16_t* label = 0;
  int16_t* o_label = 0; uint32_t* o_winner = 0; uint8_t *o_single = 0, *o_multi = 0;
  float* o_bvm = 0;                        // batch_cloud_manip: [frames][201][201] bird-view maps
  uint8_t* raw = 0; size_t raw_cap = 0;   // interleaved records of a packed batch (de-interleaved on the GPU)
  bool ensure_raw(size_t bytes) {
    if (raw_cap >= bytes) return true;
    bevgen_host_free(raw); raw_cap = bytes + bytes / 4 + 4096; raw = (uint8_t*)bevgen_host_alloc(raw_cap);
    if (!raw) raw_cap = 0;
    return raw != nullptr;
  }
  // SoA inputs: only batches that were parsed on the host need them (packed batches stage their records in `raw`), and pinning
  // memory is slow (a 16-frame HDL_64E set: 52 MB of SoA inputs, tens of milliseconds), so they are allocated on first use
  bool ensure_soa() {
    if (x) return true;
    auto A = [](size_t n) { return bevgen_host_alloc(n); };
    const size_t pts = cap_pts;
    x = (float*)A(pts * 4); y = (float*)A(pts * 4); z = (float*)A(pts * 4); inten = (float*)A(pts * 4);
    row = (uint16_t*)A(pts * 2); col = (uint16_t*)A(pts * 2); label = (int16_t*)A(pts * 2);
    return x && y && z && inten && row && col && label;
  }
  bool alloc(size_t pts, int frames, size_t S_, bool with_bvm = false) {
    release(); cap_pts = pts; cap_frames = frames; S = S_;
    auto A = [](size_t n) { return bevgen_host_alloc(n); };
    o_label = (int16_t*)A((size_t)frames * S * 2); o_winner = (uint32_t*)A(bevgen_winner_words((int64_t)pts, frames) * 4);
    o_single = (uint8_t*)A((size_t)frames * BEVGEN_GRID_SIZE * BEVGEN_GRID_SIZE);
    o_multi = (uint8_t*)A((size_t)frames * BEVGEN_NUM_LAYERS * BEVGEN_GRID_SIZE * BEVGEN_GRID_SIZE);
    if (with_bvm) o_bvm = (float*)A((size_t)frames * BEVGEN_MANIP_GRID * BEVGEN_MANIP_GRID * sizeof(float));
    return o_label && o_winner && o_single && o_multi && (!with_bvm || o_bvm);
  }
  void release_all() { release(); bevgen_host_free(raw); raw = 0; raw_cap = 0; }
  void release() {
    for (void* p : {(void*)x, (void*)y, (void*)z, (void*)inten, (void*)row, (void*)col, (void*)label, (void*)o_label, (void*)o_winner, (void*)o_single, (void*)o_multi, (void*)o_bvm}) bevgen_host_free(p);
    x = y = z = inten = 0; row = col = 0; label = 0; o_label = 0; o_winner = 0; o_single = o_multi = 0; o_bvm = 0;
  }
};

struct Batch {
  int first = 0, count = 0;
  std::vector<int64_t> offs;       // point offsets of the batch's frames (winner words are addressed through them)
  std::vector<pcdio::Cloud> clouds;
  // frames whose file is a binary PCD in the point type's own scalar types keep the file bytes instead of a parsed
  // cloud: the payload goes to the GPU as it is
  std::vector<std::vector<uint8_t>> files; std::vector<pcdio::PackedLayout> lays; std::vector<size_t> payload_pos, npts;
  std::vector<char> is_packed; bool packed = false;
  std::vector<std::string> names;
  std::vector<std::future<void>> loads;
  PinnedSet* pin = nullptr;
  std::atomic<int> pending_encodes{0};
};

struct PhaseTimer {     // adds the scope's duration to an accumulator
  std::atomic<long long>& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit PhaseTimer(std::atomic<long long>& a) : acc(a) {}
  ~PhaseTimer() { acc += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

struct Shared {
  Options opt; Dirs dirs; bevgen_params params; size_t S = 0;
  std::vector<std::string> files;
  ThreadPool* pool = nullptr;
  std::atomic<int> next_batch{0}; int n_batches = 0;
  std::mutex print_mu;
  std::atomic<long> frames_done{0};
  // CPU seconds spent per phase, summed over all pool threads (--json-metrics): where the host side of the pipeline goes
  std::atomic<long long> ns_load{0}, ns_stage{0}, ns_gpu_call{0}, ns_png{0}, ns_csv{0}, ns_bin{0}, ns_pcd{0};
  // wall time of the GPU worker threads' own (serial) sections, summed over workers: waiting for a batch's loads, for a free
  // pinned set, handing the encode tasks to the pool (--json-metrics "worker_ms_per_batch", with staging and the GPU call)
  std::atomic<long long> ns_w_loads{0}, ns_w_pin{0}, ns_w_submit{0}; std::atomic<long> batches_done{0};
  // first batch back from the GPU: everything before it is start-up (pinned / device allocations, first PCD loads)
  std::atomic<bool> first_seen{false}; std::chrono::steady_clock::time_point t_first; std::atomic<long> first_frames{0};
  std::atomic<bool> failed{false};
};

static void encode_frame(Shared& sh, const Batch& b, int k) {
  const size_t S = sh.S; const int G = BEVGEN_GRID_SIZE, L = BEVGEN_NUM_LAYERS;
  const PinnedSet& p = *b.pin;
  const std::string& name = b.names[k];
  const uint8_t* multi = p.o_multi + (size_t)k * L * G * G;
  const uint8_t* single = p.o_single + (size_t)k * G * G;
  if (sh.opt.bvm_mode) {
    if (sh.opt.encode) {   // saveAsMat, BatchCloudManip.cpp:223-238: FMT_CSV with set32fPrecision(4); imwrite converts CV_32F to 8 bit
      const int M = BEVGEN_MANIP_GRID;
      const float* g = p.o_bvm + (size_t)k * M * M;
      std::string txt = imgio::format_csv_f32(g, M, M, 4);
      std::string csv = sh.dirs.bvm + name + ".csv";
      if (!imgio::write_bytes(csv, txt.data(), txt.size())) std::cerr << "Can not open file: " << csv << "\n";
      std::vector<uint8_t> u8((size_t)M * M);
      for (int i = 0; i < M * M; i++) u8[i] = imgio::f32_to_u8_sat(g[i]);
      imgio::write_png_gray8(sh.dirs.bvm + name + ".png", u8.data(), M, M, sh.opt.png_level);
    }
  } else if (sh.opt.encode) {
    // .bin: 24 layers concatenated row-major (:294-314)
    {
      PhaseTimer t(sh.ns_bin);
      std::string bin = sh.dirs.multi_bin + name + ".bin";
      if (!imgio::write_bytes(bin, multi, (size_t)L * G * G)) std::cerr << "Can not open file: " << bin << "\n";
    }
    {
      PhaseTimer t(sh.ns_png);
      std::string img_dir = sh.dirs.multi_img + name + "/";
      mkdir(img_dir.c_str(), 0777);                                         // mkdir(2), not system("mkdir -p") (:303-306)
      char nm[16];
      for (int l = 0; l < L; l++) {
        snprintf(nm, sizeof nm, "%02d.png", l);                             // "{:02d}.png" (:316-318)
        imgio::write_png_gray8(img_dir + nm, multi + (size_t)l * G * G, G, G, sh.opt.png_level);
      }
      imgio::write_png_gray8(sh.dirs.single_img + name + ".png", single, G, G, sh.opt.png_level);   // :359-361
    }
    {
      PhaseTimer t(sh.ns_csv);
      std::string csv = sh.dirs.single_csv + name + ".csv";
      std::string txt = imgio::format_csv_u8(single, G, G);                 // :371
      if (!imgio::write_bytes(csv, txt.data(), txt.size())) std::cerr << "Faied to export csv formatted BEV file: " << csv;
    }
  }
  if (sh.opt.write_pcd) {
    PhaseTimer t(sh.ns_pcd);
    // savePCDFileBinary(non_ground/<name>.pcd, cloud_ordered) (:755-756): S slots; slot record = winning input record
    // with its label replaced by the post-ground label, empty slots all-zero.  The library reports one winner bit per
    // input point (the last writer of each (row, col) slot, :102-116); the slot is the point's own row*H + col.
    const uint32_t* win = p.o_winner + (size_t)(b.offs[k] >> 5) + (size_t)k; const int16_t* lab = p.o_label + (size_t)k * S;
    const size_t H = (size_t)sh.params.horizon_scan;
    std::string h = pcdio::header(S);
    std::vector<uint8_t> out(h.size() + S * 26, 0);
    memcpy(out.data(), h.data(), h.size());
    uint8_t* rec = out.data() + h.size();
    static const pcdio::PackedLayout canon = [] { pcdio::PackedLayout c; c.stride = 26; const int o[8] = {0, 4, 8, 12, 16, 18, 20, 24}; memcpy(c.off, o, sizeof o); return c; }();
    if (b.packed && b.lays[k] == canon) {   // the file's records ARE output records: copy 26 bytes, replace the label
      const uint8_t* pay = b.files[k].data() + b.payload_pos[k];
      for (size_t i = 0, n = b.npts[k]; i < n; i++) {
        if (!((win[i >> 5] >> (i & 31)) & 1u)) continue;
        const uint8_t* src = pay + i * 26;
        uint16_t r, c; memcpy(&r, src + 16, 2); memcpy(&c, src + 18, 2);
        const size_t s = (size_t)r * H + c;
        memcpy(rec + s * 26, src, 24); memcpy(rec + s * 26 + 24, &lab[s], 2);
      }
    } else if (b.packed) {   // winners straight from the file's own records
      const uint8_t* pay = b.files[k].data() + b.payload_pos[k]; const pcdio::PackedLayout& L = b.lays[k];
      for (size_t i = 0, n = b.npts[k]; i < n; i++) {
        if (!((win[i >> 5] >> (i & 31)) & 1u)) continue;
        float x, y, z, it; uint16_t r, c; uint32_t t; int16_t l;
        pcdio::packed_get(pay, L, i, x, y, z, it, r, c, t, l);
        const size_t s = (size_t)r * H + c;
        pcdio::pack_record(rec + s * 26, x, y, z, it, r, c, t, lab[s]);
      }
    } else {
      const pcdio::Cloud& c = b.clouds[k];
      for (size_t i = 0, n = c.size(); i < n; i++) {
        if (!((win[i >> 5] >> (i & 31)) & 1u)) continue;
        const size_t s = (size_t)c.row[i] * H + c.col[i];
        pcdio::pack_record(rec + s * 26, c.x[i], c.y[i], c.z[i], c.intensity[i], c.row[i], c.col[i], c.t[i], lab[s]);
      }
    }
    if (!pcdio::write_file(sh.dirs.non_ground + name + ".pcd", out)) std::cerr << "Can not open file: " << sh.dirs.non_ground + name + ".pcd" << "\n";
  }
}

constexpr int STAGE_THREADS = 4;
struct GpuWorker {
  Shared& sh; int dev; bevgen_ctx* ctx = nullptr; int ctx_max_pts = 0;
  std::vector<PinnedSet> pins; std::vector<PinnedSet*> free_pins; std::mutex pin_mu; std::condition_variable pin_cv;
  GpuWorker(Shared& s, int d) : sh(s), dev(d) {}

  bool ensure_ctx(int max_pts) {
    if (ctx && max_pts <= ctx_max_pts) return true;
    if (ctx) bevgen_destroy(ctx);
    ctx = nullptr;
    ctx_max_pts = std::max<int>(max_pts + max_pts / 8, (int)sh.S + 4096);
    if (bevgen_create(&ctx, dev, &sh.params, ctx_max_pts, sh.opt.batch) != 0) {
      std::cerr << "bevgen_create(device " << dev << "): " << bevgen_last_error() << std::endl;
      return false;
    }
    return true;
  }

  std::shared_ptr<Batch> grab() {
    int b = sh.next_batch++;
    if (b >= sh.n_batches) return nullptr;
    auto bt = std::make_shared<Batch>();
    bt->first = b * sh.opt.batch;
    bt->count = std::min<int>(sh.opt.batch, (int)sh.files.size() - bt->first);
    bt->clouds.resize(bt->count); bt->names.resize(bt->count);
    bt->files.resize(bt->count); bt->lays.resize(bt->count); bt->payload_pos.assign(bt->count, 0); bt->npts.assign(bt->count, 0);
    bt->is_packed.assign(bt->count, 0);
    for (int k = 0; k < bt->count; k++) {
      auto pr = std::make_shared<std::promise<void>>();
      bt->loads.push_back(pr->get_future());
      Batch* raw = bt.get(); Shared* s = &sh;
      sh.pool->submit([raw, s, k, pr, keep = bt] {
        PhaseTimer t(s->ns_load);
        const std::string& f = s->files[raw->first + k];
        raw->names[k] = short_name_of(f);
        std::string err;
        bool done = false;
        if (s->opt.packed) {
          pcdio::Header h;
          if (pcdio::read_file(f, raw->files[k], &err) && pcdio::parse_header(raw->files[k], h)) {
            if (pcdio::packed_layout(h, raw->lays[k])) {
              raw->payload_pos[k] = h.payload_pos;
              raw->npts[k] = std::min(h.n, (raw->files[k].size() - h.payload_pos) / (size_t)h.rec);
              raw->is_packed[k] = 1; done = true;
            } else {
              done = pcdio::decode(raw->files[k], h, f, raw->clouds[k], &err);
              raw->files[k].clear(); raw->files[k].shrink_to_fit();
              if (!done) { std::cerr << "[pcd] " << err << std::endl; done = true; }
            }
          }
        }
        if (!done && !pcdio::load(f, raw->clouds[k], &err)) std::cerr << "[pcd] " << err << std::endl;   // reference ignores the status (:730)
        pr->set_value();
      });
    }
    return bt;
  }

  PinnedSet* take_pin(size_t pts) {
    std::unique_lock<std::mutex> l(pin_mu);
    pin_cv.wait(l, [&] { return !free_pins.empty(); });
    PinnedSet* p = free_pins.back(); free_pins.pop_back();
    l.unlock();
    if (p->cap_pts < pts || p->cap_frames < sh.opt.batch) {
      if (!p->alloc(pts + pts / 4 + 1024, sh.opt.batch, sh.S, sh.opt.bvm_mode)) { std::cerr << "pinned allocation failed" << std::endl; sh.failed = true; }
    }
    return p;
  }
  void give_pin(PinnedSet* p) { { std::lock_guard<std::mutex> l(pin_mu); free_pins.push_back(p); } pin_cv.notify_one(); }

  void run() {
    // Staging threads of this worker alone: the shared pool's threads sit in encode tasks of several milliseconds, so helpers
    // queued there arrived late and the worker copied most of a batch by itself (measured: 11 ms per 16-frame batch, 4.5 GB/s,
    // the largest serial section of the pipeline)
    ThreadPool stage_pool(STAGE_THREADS);
    pins.resize(4);
    for (auto& p : pins) free_pins.push_back(&p);
    // Three batches of PCD loads are always queued ahead of the batch on the GPU: the pool is FIFO, so a batch's loads sit in
    // front of the encode tasks of the batches before it and the GPU never waits for a file behind a queue of PNG / PCD writers
    std::deque<std::shared_ptr<Batch>> ahead;
    auto refill = [&] { while (ahead.size() < 3) { auto b = grab(); if (!b) break; ahead.push_back(b); } };
    refill();
    while (!ahead.empty() && !sh.failed) {
      std::shared_ptr<Batch> cur = ahead.front(); ahead.pop_front();
      refill();
      { PhaseTimer t(sh.ns_w_loads); for (auto& f : cur->loads) f.get(); }
      // the batch goes through the packed path iff every frame is an interleaved payload of one and the same layout
      cur->packed = sh.opt.packed && cur->count > 0;
      for (int k = 0; k < cur->count && cur->packed; k++) cur->packed = cur->is_packed[k] && cur->lays[k] == cur->lays[0];
      if (!cur->packed)
        for (int k = 0; k < cur->count; k++)
          if (cur->is_packed[k]) {   // mixed batch: parse this frame on the host after all
            pcdio::Header h; std::string err;
            if (!pcdio::parse_header(cur->files[k], h) || !pcdio::decode(cur->files[k], h, sh.files[cur->first + k], cur->clouds[k], &err)) std::cerr << "[pcd] " << err << std::endl;
            cur->is_packed[k] = 0; cur->files[k].clear(); cur->files[k].shrink_to_fit();
          }
      auto frame_n = [&](int k) { return cur->packed ? cur->npts[k] : cur->clouds[k].size(); };
      std::vector<int64_t> offs(cur->count + 1, 0);
      int max_n = 0;
      for (int k = 0; k < cur->count; k++) { offs[k + 1] = offs[k] + (int64_t)frame_n(k); max_n = std::max<int>(max_n, (int)frame_n(k)); }
      if (!ensure_ctx(max_n)) { sh.failed = true; break; }
      PinnedSet* p;
      { PhaseTimer t(sh.ns_w_pin); p = take_pin((size_t)offs[cur->count]); }
      if (sh.failed) break;
      cur->pin = p;
      { std::lock_guard<std::mutex> l(sh.print_mu); for (int k = 0; k < cur->count; k++) std::cout << "Converting file: " << cur->names[k] << "\n"; }   // :744
      bevgen_outputs out{p->o_label, p->o_winner, p->o_single, p->o_multi, sh.opt.bvm_mode ? p->o_bvm : nullptr};
      cur->offs = offs;
      int rc;
      const auto t_stage0 = Clock::now();
      if (cur->packed) {
        const size_t stride = (size_t)cur->lays[0].stride;
        if (!p->ensure_raw((size_t)offs[cur->count] * stride + 64)) { std::cerr << "pinned allocation failed" << std::endl; sh.failed = true; give_pin(p); break; }
        stage_pool.parallel_for(cur->count, [&](int k) {  // staging = one copy of the file payload into pinned memory, frames in parallel
          if (cur->npts[k]) memcpy(p->raw + (size_t)offs[k] * stride, cur->files[k].data() + cur->payload_pos[k], cur->npts[k] * stride);
        });
        const int* o = cur->lays[0].off;
        bevgen_record_layout lay{cur->lays[0].stride, o[0], o[1], o[2], o[3], o[4], o[5], o[7]};
        sh.ns_stage += std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t_stage0).count();
        PhaseTimer t(sh.ns_gpu_call);
        rc = bevgen_process_packed_host(ctx, cur->count, offs.data(), p->raw, &lay, &out);
      } else {
        if (!p->ensure_soa()) { std::cerr << "pinned allocation failed" << std::endl; sh.failed = true; give_pin(p); break; }
        stage_pool.parallel_for(cur->count, [&](int k) {    // SoA staging into pinned memory, frames in parallel
          const pcdio::Cloud& c = cur->clouds[k]; size_t o = (size_t)offs[k], n = c.size();
          if (!n) return;
          memcpy(p->x + o, c.x.data(), n * 4); memcpy(p->y + o, c.y.data(), n * 4); memcpy(p->z + o, c.z.data(), n * 4);
          memcpy(p->inten + o, c.intensity.data(), n * 4); memcpy(p->row + o, c.row.data(), n * 2); memcpy(p->col + o, c.col.data(), n * 2);
          memcpy(p->label + o, c.label.data(), n * 2);
        });
        bevgen_points in{p->x, p->y, p->z, p->inten, p->row, p->col, p->label};
        sh.ns_stage += std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t_stage0).count();
        PhaseTimer t(sh.ns_gpu_call);
        rc = bevgen_process_host(ctx, cur->count, offs.data(), &in, &out);
      }
      if (rc != 0) {
        std::cerr << "bevgen_process_host: " << bevgen_last_error() << std::endl; sh.failed = true; give_pin(p); break;
      }
      if (!sh.first_seen.exchange(true)) { sh.t_first = std::chrono::steady_clock::now(); sh.first_frames = cur->count; }
      cur->pending_encodes = cur->count;
      {
        PhaseTimer t(sh.ns_w_submit);
        for (int k = 0; k < cur->count; k++) {
          sh.pool->submit([this, cur, k] {
            encode_frame(sh, *cur, k);
            sh.frames_done++;
            if (--cur->pending_encodes == 0) give_pin(cur->pin);
          });
        }
      }
      sh.batches_done++;
    }
  }
};

int main(int argc, char** argv) {
  // Every frame allocates and frees a few 3.4 MB buffers (file bytes, output PCD) on the pool's threads.  glibc serves such sizes
  // with mmap / munmap: 830 page faults per buffer and an address-space lock all threads share.  Keeping freed memory in the heap
  // instead (it stops growing at the pipeline's depth, a few hundred MB) measured 1377 -> 1580 frames/s after the first batch on a
  // 1200-keyframe folder (PCD read 1.9 -> 1.4 ms, PCD write 5.3 -> 4.4 ms per frame; profiles/r2_cli_probe_malloc_threads.log).
  mallopt(M_MMAP_THRESHOLD, 256 << 20); mallopt(M_TRIM_THRESHOLD, 0x7fffffff); mallopt(M_TOP_PAD, 64 << 20);
#ifdef BATCH_CLOUD_MANIP
  if (argc < 2 || argv[1] == nullptr) { std::cout << "Usage: " << argv[0] << " <keyframes_root_dir>" << std::endl; exit(1); }   // BatchCloudManip.cpp:271-274
  Shared sh;
  Options& opt = sh.opt;
  opt.root = argv[1]; opt.sensor = "HDL_64E"; opt.bvm_mode = true;   // N_SCAN 64, Horizon_SCAN 2083, groundScanInd 50 (:12-13, :85)
  const int first_opt = 2;
#else
  if (argc < 3 || argv[1] == nullptr || argv[2] == nullptr) usage_and_exit(argv[0]);
  Shared sh;
  Options& opt = sh.opt;
  opt.root = argv[1]; opt.sensor = argv[2];
  const int first_opt = 3;
#endif
  for (int i = first_opt; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&](const char* name) -> const char* { if (i + 1 >= argc) { std::cerr << name << " needs a value\n"; exit(1); } return argv[++i]; };
    if (a == "--gpus") opt.gpus = atoi(val("--gpus"));
    else if (a == "--batch") opt.batch = atoi(val("--batch"));
    else if (a == "--threads") opt.threads = atoi(val("--threads"));
    else if (a == "--png-level") opt.png_level = atoi(val("--png-level"));
    else if (a == "--workers-per-gpu") opt.workers_per_gpu = std::max(1, std::min(8, atoi(val("--workers-per-gpu"))));
    else if (a == "--no-encode") opt.encode = false;
    else if (a == "--no-pcd") opt.write_pcd = false;
    else if (a == "--no-packed") opt.packed = false;
    else if (a == "--json-metrics") opt.json_metrics = val("--json-metrics");
    else { std::cerr << "unknown option " << a << "\n"; exit(1); }
  }
  if (opt.gpus < 1) opt.gpus = 1;
  if (opt.batch < 1) opt.batch = 1;
  if (opt.threads <= 0) opt.threads = std::max(2u, std::thread::hardware_concurrency());
  std::string root = opt.root;
  if (root.back() != '/') root.append("/");                                   // :691-693
  Dirs& d = sh.dirs;
  d.pcd_in = root + "keyframe_point_cloud/"; d.non_ground = root + "non_ground_point_cloud/";
  d.pose_file = root + "keyframe_pose.csv"; d.label_file = root + "keyframe_label.csv";
  d.multi_bin = root + "output_multi_bev/binary/"; d.multi_img = root + "output_multi_bev/image/";
  d.single_csv = root + "output_single_bev/csv/"; d.single_img = root + "output_single_bev/image/";

  d.bvm = root + "output_bvm/";
  reset_dir(d.non_ground);                                                    // :704-705
  sh.files = list_pcd(d.pcd_in);                                              // :708-709
  if (opt.bvm_mode) reset_dir(d.bvm);                                         // BatchCloudManip.cpp:291-295
  else {
    reset_dir(root + "output_multi_bev/"); reset_dir(d.multi_bin); reset_dir(d.multi_img);   // initDirectories :39-71
    reset_dir(d.single_csv); reset_dir(d.single_img);
  }

  if (bevgen_sensor_params(opt.sensor.c_str(), &sh.params) < 0) {             // :718-719
    std::cerr << "Unknown sensor type: " << opt.sensor << "!" << std::endl;
    std::cerr << "Unknown sensor type! " << std::endl;
    return 1;   // the reference would go on with uninitialised SensorParams (undefined behaviour); we stop
  }
  sh.S = (size_t)sh.params.n_scan * sh.params.horizon_scan;
  if (!opt.bvm_mode)
    std::cout << "Using sensor_type " << opt.sensor << ", with params: N_SCAN: " << sh.params.n_scan << ", Horizon_SCAN: "
              << sh.params.horizon_scan << ", GROUND_UPPER_SCAN: " << sh.params.ground_upper_scan << "\n";   // :720-722

  ThreadPool pool(opt.threads);
  sh.pool = &pool;
  sh.n_batches = (int)((sh.files.size() + opt.batch - 1) / opt.batch);
  const int n_gpus = std::max(1, std::min(opt.gpus, std::max(1, sh.n_batches)));
  const int n_workers = std::max(n_gpus, std::min(n_gpus * opt.workers_per_gpu, std::max(1, sh.n_batches)));
  std::vector<std::unique_ptr<GpuWorker>> workers;
  for (int g = 0; g < n_workers; g++) workers.emplace_back(new GpuWorker(sh, g % n_gpus));
  // contexts are also needed for the label stage even when there is no frame to process
  for (auto& w : workers) if (!w->ensure_ctx((int)sh.S)) return 1;

  auto t0 = Clock::now();
  {
    std::vector<std::thread> ths;
    for (auto& w : workers) ths.emplace_back([&w] { w->run(); });
    for (auto& t : ths) t.join();
    pool.wait_idle();
  }
  const auto t_end = Clock::now();
  double total_ms = std::chrono::duration<double, std::milli>(t_end - t0).count();
  // start-up (allocations + the first batch's loads and GPU pass) and the rate of the pipeline once it is running
  double startup_ms = 0.0, steady_fps = 0.0;
  if (sh.first_seen) {
    startup_ms = std::chrono::duration<double, std::milli>(sh.t_first - t0).count();
    const double rest_s = std::chrono::duration<double>(t_end - sh.t_first).count();
    if (rest_s > 0 && (long)sh.files.size() > sh.first_frames) steady_fps = ((long)sh.files.size() - sh.first_frames) / rest_s;
  }
  if (sh.failed) return 1;
  // the reference averages its per-frame serial span (:749-759); here frames overlap, so this is wall time / frames
  std::cout << "[TIME] Average preprocessing and BEV generation: " << (sh.files.empty() ? 0.0 : total_ms / sh.files.size()) << "\n";

  if (opt.bvm_mode) {   // batch_cloud_manip has no label stage (BatchCloudManip.cpp:327-330)
    if (!opt.json_metrics.empty()) {
      std::ofstream j(opt.json_metrics);
      j << "{\"frames\": " << sh.files.size() << ", \"gpus\": " << n_gpus << ", \"frames_wall_ms\": " << total_ms << "}\n";
    }
    for (auto& wk : workers) { if (wk->ctx) bevgen_destroy(wk->ctx); for (auto& p : wk->pins) p.release_all(); }
    std::cout << "Done. " << std::endl;
    return 0;
  }
  // Step 2: labels (:761-765)
  auto t1 = Clock::now();
  std::vector<float> xyz = read_poses(d.pose_file);
  const int K = (int)(xyz.size() / 3);
  std::vector<int32_t> major(std::max(K, 1)), overlap(std::max(K, 1));
  int32_t M = 0;
  if (K > 0) {
    if (bevgen_select_major(workers[0]->ctx, K, xyz.data(), major.data(), &M, overlap.data()) != 0) {
      std::cerr << "bevgen_select_major: " << bevgen_last_error() << std::endl; return 1;
    }
    for (int i = 1; i < K; i++)
      if (overlap[i] >= 0)                                                      // :553-555
        std::cout << "Key Frame " << i << " overlaps with previous Major Frame " << overlap[i] << ", i.e. Key Frame " << major[overlap[i]] << ". \n";
  }
  std::cout << "One-hot label has length: " << M << std::endl;               // :580-581
  std::vector<int32_t> nn((size_t)std::max(K, 1) * 2); std::vector<float> w((size_t)std::max(K, 1) * 2);
  if (K > 0) {
    // rows split per GPU, gathered in host memory
    const int nw = (int)workers.size();
    std::vector<std::thread> ths; std::atomic<bool> bad{false};
    for (int g = 0; g < nw; g++) {
      const int r0 = (int)((int64_t)K * g / nw), r1 = (int)((int64_t)K * (g + 1) / nw);
      if (r0 == r1) continue;
      ths.emplace_back([&, g, r0, r1] {
        if (bevgen_labels(workers[g]->ctx, K, xyz.data(), M, major.data(), r0, r1, nullptr, nn.data() + 2 * (size_t)r0, w.data() + 2 * (size_t)r0) != 0) {
          std::cerr << "bevgen_labels: " << bevgen_last_error() << std::endl; bad = true;
        }
      });
    }
    for (auto& t : ths) t.join();
    if (bad) return 1;
  }
  {
    // saveLabels (:645-661): every value through `ostream << float` followed by ',', rows end with '\n'
    std::ofstream f(d.label_file);
    if (!f.is_open()) { std::cerr << "failed to open keyframe label file: " << d.label_file << std::endl; exit(1); }
    std::string zero_row; for (int j = 0; j < M; j++) zero_row += "0,";
    for (int i = 0; i < K; i++) {
      // a row has at most two non-zeros; emit runs of "0," around them (identical text to streaming M floats)
      int a = nn[2 * i], b = nn[2 * i + 1]; float wa = w[2 * i], wb = w[2 * i + 1];
      std::vector<std::pair<int, float>> nz;
      if (b >= 0 && b == a) nz = {{a, wb}};                                      // M == 1 edge: w1 overwrote w0 (:629-630)
      else { nz.push_back({a, wa}); if (b >= 0) nz.push_back({b, wb}); }
      std::sort(nz.begin(), nz.end());
      int col = 0; std::ostringstream row;
      for (auto& e : nz) { row.write(zero_row.data(), 2 * (size_t)(e.first - col)); row << e.second << ","; col = e.first + 1; }
      row.write(zero_row.data(), 2 * (size_t)(M - col));
      f << row.str() << "\n";
    }
    std::cout << "saved labels from " << K << " key frames. " << std::endl;   // :659-660
  }
  double label_ms = std::chrono::duration<double, std::milli>(Clock::now() - t1).count();
  if (!opt.json_metrics.empty()) {
    std::ofstream j(opt.json_metrics);
    const double nfr = std::max<double>(1.0, (double)sh.files.size());
    j << "{\"frames\": " << sh.files.size() << ", \"gpus\": " << n_gpus << ", \"workers\": " << workers.size() << ", \"batch\": " << opt.batch << ", \"threads\": " << opt.threads
      << ", \"encode\": " << (opt.encode ? "true" : "false") << ", \"write_pcd\": " << (opt.write_pcd ? "true" : "false")
      << ", \"frames_wall_ms\": " << total_ms << ", \"frames_per_s\": " << (total_ms > 0 ? sh.files.size() / (total_ms * 1e-3) : 0.0)
      << ", \"startup_ms\": " << startup_ms << ", \"frames_per_s_after_first_batch\": " << steady_fps
      << ", \"cpu_ms_per_frame\": {\"pcd_load\": " << sh.ns_load * 1e-6 / nfr << ", \"pinned_staging\": " << sh.ns_stage * 1e-6 / nfr
      << ", \"gpu_call_wall\": " << sh.ns_gpu_call * 1e-6 / nfr << ", \"bin\": " << sh.ns_bin * 1e-6 / nfr << ", \"png_25\": " << sh.ns_png * 1e-6 / nfr
      << ", \"csv\": " << sh.ns_csv * 1e-6 / nfr << ", \"pcd_write\": " << sh.ns_pcd * 1e-6 / nfr << "}"
      << ", \"worker_ms_per_batch\": {\"batches\": " << sh.batches_done << ", \"wait_loads\": " << sh.ns_w_loads * 1e-6 / std::max<double>(1.0, (double)sh.batches_done)
      << ", \"wait_pinned_set\": " << sh.ns_w_pin * 1e-6 / std::max<double>(1.0, (double)sh.batches_done)
      << ", \"staging\": " << sh.ns_stage * 1e-6 / std::max<double>(1.0, (double)sh.batches_done)
      << ", \"gpu_call\": " << sh.ns_gpu_call * 1e-6 / std::max<double>(1.0, (double)sh.batches_done)
      << ", \"submit_encodes\": " << sh.ns_w_submit * 1e-6 / std::max<double>(1.0, (double)sh.batches_done) << "}"
      << ", \"keyframes\": " << K << ", \"majors\": " << M << ", \"labels_wall_ms\": " << label_ms << "}\n";
  }
  for (auto& wk : workers) { if (wk->ctx) bevgen_destroy(wk->ctx); for (auto& p : wk->pins) p.release_all(); }
  std::cout << "Done. " << std::endl;
  return 0;
}
