// Host-side I/O pipeline of spimFusionBatch: read-ahead of the next time point's two TIFF stacks and
// write-behind of the output stacks, so that disk I/O and the 16-bit <-> float conversions overlap
// the GPU work of the current time point.  The reference reads, computes and writes strictly in
// sequence (src/spim_fusion_batch.cpp:668-940); files, names and contents are unchanged here.
// MILB_PIPELINE=0 restores the sequential behaviour.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "cli_common.h"

inline bool pipeline_enabled()
{
	const char *e = getenv("MILB_PIPELINE");
	return !(e && e[0] == '0');
}

// Reads (path1, path2) in the background; take() hands the stacks over if they are the ones asked for.
class ReadAhead {
public:
	~ReadAhead() { join(); }
	// raw16: keep 16-bit stacks as they are on disk (the device-resident pipeline converts them to float on the GPU)
	void start(const std::string &p1, const std::string &p2, size_t n1, size_t n2, bool raw16 = false)
	{
		join();
		if (!pipeline_enabled() || !fexists((char *)p1.c_str()) || !fexists((char *)p2.c_str())) return;
		k1_ = p1; k2_ = p2;
		raw16_ = raw16;
		if (raw16) { u1_.resize(n1); u2_.resize(n2); }
		else { b1_.resize(n1); b2_.resize(n2); }
		busy_ = true;
		th_ = std::thread([this] { // the two views are decoded side by side
			std::thread second([this] {
				if (raw16_) read_stack_checked(u2_.data(), k2_, u2_.size(), s2_);
				else read_stack_checked(b2_.data(), k2_, b2_.size(), s2_);
			});
			if (raw16_) read_stack_checked(u1_.data(), k1_, u1_.size(), s1_);
			else read_stack_checked(b1_.data(), k1_, b1_.size(), s1_);
			second.join();
		});
	}
	// 16-bit variant of take()
	bool take16(const std::string &p1, const std::string &p2, HostVec16 &raw1, HostVec16 &raw2, unsigned int *s1, unsigned int *s2)
	{
		if (!busy_) return false;
		join();
		if (p1 != k1_ || p2 != k2_ || !raw16_) return false;
		raw1.swap(u1_); raw2.swap(u2_);
		memcpy(s1, s1_, sizeof s1_); memcpy(s2, s2_, sizeof s2_);
		return true;
	}
	// true: raw1/raw2 and the size triples now hold the prefetched stacks
	bool take(const std::string &p1, const std::string &p2, HostVec &raw1, HostVec &raw2, unsigned int *s1, unsigned int *s2)
	{
		if (!busy_) return false;
		join();
		if (p1 != k1_ || p2 != k2_ || raw16_) return false;
		raw1.swap(b1_); raw2.swap(b2_);
		memcpy(s1, s1_, sizeof s1_); memcpy(s2, s2_, sizeof s2_);
		return true;
	}

private:
	void join()
	{
		if (th_.joinable()) th_.join();
		busy_ = false;
	}
	std::thread th_;
	bool busy_ = false;
	std::string k1_, k2_;
	HostVec b1_, b2_;
	HostVec16 u1_, u2_;
	bool raw16_ = false;
	unsigned int s1_[3] = {0, 0, 0}, s2_[3] = {0, 0, 0};
};

// writetifstack on a worker thread; the data is copied at submission so the caller can reuse its buffer.
class WriteBehind {
public:
	WriteBehind()
	{
		if (!pipeline_enabled()) return;
		int n = 3; // encoder / writer threads: one 16-bit conversion + fwrite stream does not keep up with the GPU
		if (const char *e = getenv("MILB_WRITERS")) n = atoi(e);
		for (int i = 0; i < (n < 1 ? 1 : n); i++) th_.emplace_back([this] { run(); });
	}
	~WriteBehind() { drain(); }
	void write(const std::string &path, const float *data, const unsigned int *size, unsigned short bits)
	{
		if (th_.empty()) { // sequential mode
			unsigned int s[3] = {size[0], size[1], size[2]};
			writetifstack((char *)path.c_str(), (float *)data, s, bits);
			return;
		}
		Job j;
		j.path = path; j.bits = bits;
		memcpy(j.size, size, sizeof j.size);
		j.data.assign(data, data + voxels(size));
		std::unique_lock<std::mutex> lk(mu_);
		cv_space_.wait(lk, [this] { return q_.size() < 6; }); // bound the host memory held by pending writes
		q_.push_back(std::move(j));
		cv_work_.notify_one();
	}
	// like write(), but takes the caller's buffer instead of copying it: `buf` comes back holding a recycled
	// buffer of the same size with unspecified contents (the caller overwrites it for the next time point)
	void write_swap(const std::string &path, HostVec &buf, const unsigned int *size, unsigned short bits)
	{
		if (th_.empty()) { write(path, buf.data(), size, bits); return; }
		Job j;
		j.path = path; j.bits = bits;
		memcpy(j.size, size, sizeof j.size);
		{
			std::unique_lock<std::mutex> lk(mu_);
			if (!spare_.empty()) { j.big.swap(spare_.back()); spare_.pop_back(); }
		}
		if (j.big.size() != buf.size()) j.big.resize(buf.size());
		j.big.swap(buf);
		std::unique_lock<std::mutex> lk(mu_);
		cv_space_.wait(lk, [this] { return q_.size() < 6; });
		q_.push_back(std::move(j));
		cv_work_.notify_one();
	}
	// a 16-bit stack that is already converted (on the GPU): the buffer is taken over like in write_swap and written as is
	void write_u16_swap(const std::string &path, HostVec16 &buf, const unsigned int *size)
	{
		if (th_.empty()) {
			unsigned int s[3] = {size[0], size[1], size[2]};
			writetifstack_16to16((char *)path.c_str(), buf.data(), s);
			return;
		}
		Job j;
		j.path = path; j.bits = 16;
		memcpy(j.size, size, sizeof j.size);
		{
			std::unique_lock<std::mutex> lk(mu_);
			if (!spare16_.empty()) { j.big16.swap(spare16_.back()); spare16_.pop_back(); }
		}
		if (j.big16.size() != buf.size()) j.big16.resize(buf.size());
		j.big16.swap(buf);
		std::unique_lock<std::mutex> lk(mu_);
		cv_space_.wait(lk, [this] { return q_.size() < 6; });
		q_.push_back(std::move(j));
		cv_work_.notify_one();
	}
	// waits until everything submitted so far is on disk (end of the batch)
	void drain()
	{
		if (th_.empty()) return;
		{
			std::unique_lock<std::mutex> lk(mu_);
			stop_ = true;
			cv_work_.notify_all();
		}
		for (auto &t : th_) t.join();
		th_.clear();
	}

private:
	struct Job {
		std::string path;
		std::vector<float> data; // copied payload (small outputs) ...
		HostVec big;             // ... or a buffer taken over from the caller (write_swap)
		HostVec16 big16;         // ... or an already converted 16-bit stack (write_u16_swap)
		unsigned int size[3];
		unsigned short bits;
	};
	void run()
	{
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> lk(mu_);
				cv_work_.wait(lk, [this] { return stop_ || !q_.empty(); });
				if (q_.empty()) return;
				j = std::move(q_.front());
				q_.pop_front();
				cv_space_.notify_one();
			}
			if (!j.big16.empty()) {
				writetifstack_16to16((char *)j.path.c_str(), j.big16.data(), j.size);
				std::unique_lock<std::mutex> lk(mu_);
				spare16_.push_back(std::move(j.big16));
				continue;
			}
			writetifstack((char *)j.path.c_str(), j.big.empty() ? j.data.data() : j.big.data(), j.size, j.bits);
			if (!j.big.empty()) {
				std::unique_lock<std::mutex> lk(mu_);
				spare_.push_back(std::move(j.big));
			}
		}
	}
	std::vector<std::thread> th_;
	std::mutex mu_;
	std::condition_variable cv_work_, cv_space_;
	std::deque<Job> q_;
	std::vector<HostVec> spare_; // written-out big buffers waiting to be reused
	std::vector<HostVec16> spare16_;
	bool stop_ = false;
};
