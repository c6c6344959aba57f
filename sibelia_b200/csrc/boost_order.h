// Iteration order of boost::unordered_map<size_t, T> as vendored by the reference (Boost 1.54, 64-bit size_t).
//
// AnyBulges (/root/reference/src/bulgeremoval.cpp:168,203-215) collects the bulge groups of one vertex in a
// boost::unordered_map keyed by vertex id and then walks the map from begin() to end(); that order decides the order
// in which the bulges of the vertex are collapsed, so it is part of the result.  This header restates exactly the
// part of the container that determines it (not the container itself):
//   * hash:        boost::hash<size_t> is the identity; mix64_policy::apply_hash then applies Thomas Wang's 64-bit
//                  mix and the bucket is hash & (bucket_count - 1)         .../unordered/detail/buckets.hpp:604-620
//   * growth:      default-constructed map: 16 buckets on first insert (default_bucket_count = 11 rounded to a power
//                  of two, util.hpp:27, buckets.hpp:623-633), max load factor 1.0; the insert that would make
//                  size > bucket_count first rehashes to new_bucket_count(max(size+1, size + size/2) + 1)
//                                                                          .../detail/table.hpp:321-335, 808-822
//   * node order:  all nodes form one singly linked list; a node whose bucket is empty becomes the list head, a node
//                  whose bucket is occupied is linked right after the bucket's predecessor node
//                                                                          .../detail/unique.hpp:302-333
//   * rehash:      nodes are re-threaded in current list order with the same two rules     .../unique.hpp:591-618
// Differentially tested against the vendored header itself (tests/test_boost_order.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sibgpu {

class BoostUnorderedOrder {
public:
	BoostUnorderedOrder() { clear(); }

	void clear()
	{
		hash_.clear();
		next_.clear();
		payload_.clear();
		bucket_prev_.clear();
		start_next_ = NIL;
		bucket_count_ = 16;
		have_buckets_ = false;
		max_load_ = 0;
	}

	static uint64_t mix(uint64_t key)
	{
		key = (~key) + (key << 21);
		key = key ^ (key >> 24);
		key = (key + (key << 3)) + (key << 8);
		key = key ^ (key >> 14);
		key = (key + (key << 2)) + (key << 4);
		key = key ^ (key >> 28);
		key = key + (key << 31);
		return key;
	}

	// Inserts a key that is NOT present yet (the caller keeps its own key -> payload lookup); payload is an opaque
	// index handed back by order().
	void insert_new(uint64_t key, int payload)
	{
		const uint64_t h = mix(key);
		reserve_for_insert(hash_.size() + 1);
		const int n = (int)hash_.size();
		hash_.push_back(h);
		next_.push_back(NIL);
		payload_.push_back(payload);
		const size_t b = h & (bucket_count_ - 1);
		if(bucket_prev_[b] == NONE)
		{
			if(start_next_ != NIL) bucket_prev_[hash_[start_next_] & (bucket_count_ - 1)] = n;
			bucket_prev_[b] = START;
			next_[n] = start_next_;
			start_next_ = n;
		}
		else
		{
			const int pr = bucket_prev_[b];
			next_[n] = next_of(pr);
			next_of(pr) = n;
		}
	}

	// payloads in begin()..end() order
	template<class Out> void order(Out out) const
	{
		for(int n = start_next_; n != NIL; n = next_[n]) *out++ = payload_[n];
	}

	size_t size() const { return hash_.size(); }

private:
	enum { NIL = -1,               // null link
	       NONE = -1,              // empty bucket
	       START = -2 };           // the dummy start node

	int &next_of(int link) { return link == START ? start_next_ : next_[link]; }

	static size_t new_bucket_count(size_t min)
	{
		if(min <= 4) return 4;
		--min;
		min |= min >> 1; min |= min >> 2; min |= min >> 4; min |= min >> 8; min |= min >> 16; min |= min >> 32;
		return min + 1;
	}

	static size_t min_buckets_for_size(size_t size) { return new_bucket_count(size + 1); }   // mlf == 1.0

	void reserve_for_insert(size_t size)
	{
		if(!have_buckets_)
		{
			size_t n = min_buckets_for_size(size);
			if(n < bucket_count_) n = bucket_count_;
			bucket_count_ = n;
			bucket_prev_.assign(n, (int)NONE);
			max_load_ = n;
			have_buckets_ = true;
		}
		else if(size > max_load_)
		{
			const size_t cur = hash_.size();
			const size_t want = size > cur + (cur >> 1) ? size : cur + (cur >> 1);
			const size_t n = min_buckets_for_size(want);
			if(n != bucket_count_) rehash(n);
		}
	}

	void rehash(size_t n)
	{
		bucket_count_ = n;
		bucket_prev_.assign(n, (int)NONE);
		max_load_ = n;
		int prev = START;
		while(next_of(prev) != NIL)
		{
			const int node = next_of(prev);
			const size_t b = hash_[node] & (n - 1);
			if(bucket_prev_[b] == NONE)
			{
				bucket_prev_[b] = prev;
				prev = node;
			}
			else
			{
				next_of(prev) = next_[node];
				const int bp = bucket_prev_[b];
				next_[node] = next_of(bp);
				next_of(bp) = node;
			}
		}
	}

	std::vector<uint64_t> hash_;
	std::vector<int> next_, payload_, bucket_prev_;
	int start_next_;
	size_t bucket_count_, max_load_;
	bool have_buckets_;
};

} // namespace sibgpu
