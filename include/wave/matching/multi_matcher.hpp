// wave::MultiMatcher<T, R> - batch dispatch of independent matches, same surface as the reference
// (wave_matching/include/wave/matching/multi_matcher.hpp:30-96): a fixed pool of workers, each
// owning one matcher; insert() blocks while the bounded input queue is full; done() tells whether
// every inserted pair has been matched; getResult() - declared but never defined in the reference
// (multi_matcher.hpp:73) - pops the oldest finished result as its doc comment describes.
//
// On the B200 a worker is a host thread driving its own C-ABI handle, i.e. its own CUDA stream and
// device buffers: matches of different workers overlap on the GPU.  With WAVE_MATCHING_DEVICE=all the
// workers' matchers (ICP, GICP and NDT alike) are spread round-robin over the visible devices, so on
// the 8-GPU box the same class shards a batch of scans across all GPUs with no data-path
// communication; unset, every matcher sits on device 0.
#ifndef WAVE_MATCHING_MULTI_MATCHER_HPP
#define WAVE_MATCHING_MULTI_MATCHER_HPP

#include <condition_variable>
#include <exception>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "wave/matching/matcher.hpp"
#include "wave/matching/pcl_common.hpp"
#include "wave/utils/log.hpp"

namespace wave {

template <typename T, typename R>
class MultiMatcher {
 public:
    MultiMatcher(int n_threads = static_cast<int>(std::thread::hardware_concurrency()), int queue_s = 10,
                 R params = R())
        : n_thread(n_threads), queue_size(queue_s), remaining_matches(0), config(params), stop(false) {
        this->initPool(params);
    }
    ~MultiMatcher();

    /// queue a pair; blocks while the input queue already holds queue_size pairs
    void insert(const int &id, const PCLPointCloudPtr &src, const PCLPointCloudPtr &target);
    /// true when every inserted pair has been matched
    bool done();
    /// pops the oldest finished result; blocks while results are pending; false once drained
    bool getResult(int *id, Eigen::Affine3d *transform, Mat6 *info);
    /// Extension for scan-to-map batches (not in the reference; matchers with shareTarget(), i.e.
    /// ICPMatcher): every job inserted with a null target is matched against `map`, which is uploaded
    /// and indexed once per GPU instead of once per job.  Call while no job is pending.
    void setMap(const PCLPointCloudPtr &map);

 private:
    const int n_thread;
    const int queue_size;
    int remaining_matches;
    R config;
    std::queue<std::tuple<int, PCLPointCloudPtr, PCLPointCloudPtr>> input;
    std::queue<std::tuple<int, Eigen::Affine3d, Mat6>> output;
    std::vector<std::thread> pool;
    std::vector<T *> matchers;
    std::vector<T *> map_owners;   // one per device in use (setMap)

    std::mutex ip_mutex, op_mutex, cnt_mutex;
    std::condition_variable ip_condition, op_condition;
    bool stop;

    void spin(int threadid);
    void initPool(R params);
};

}  // namespace wave

#include "wave/matching/impl/multi_matcher_impl.hpp"

#endif  // WAVE_MATCHING_MULTI_MATCHER_HPP
