// RingSuballocator — suballocation of one device buffer, placement-compatible with the reference's
// CadR::CircularAllocationMemory (src/CadR/CircularAllocationMemory.h) but built differently.
//
// Behaviour that applications (and tests/DataAllocationTest.cpp) can observe, all reproduced here and pinned by
// tests/golden/allocator_kat.json.gz (vectors produced by the reference's own header):
//   * two address-ordered regions: region 1 bumps towards the end of the buffer; when a request no longer fits
//     there, region 2 bumps from the start of the buffer up to region 1's oldest live allocation
//     (CircularAllocationMemory.h:521-560);
//   * alignment 64 for requests >= 64 bytes, otherwise 16 (:564-568);
//   * space is reclaimed from the OLDEST end only: a region's start moves to its oldest live allocation, and a
//     region whose allocations are all gone is reset; when region 1 empties while region 2 holds data, region 2
//     takes over the role of region 1 (:320-347);
//   * usedBytes() counts alignment padding: +(end - previous region end) on allocation, -(next record's
//     address - own address) on release (:574-600, :619-639).
//
// Design: each region is a std::deque of records in allocation order (stable addresses under push_back/pop_front),
// a released record is only flagged, and dead records are popped from the front — no magic values, no fixed-size
// record blocks.  Zero-size "pin" records keep a region from being reclaimed past them; DataMemory uses one per
// upload run, which is what the reference's staging markers do (DataMemory.cpp:292-387).
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <utility>

namespace CadR {

struct RingRegionBase;

struct RingRecord {
	uint64_t deviceAddress = 0;   ///< start of the allocation (for a pin: where the next allocation will start)
	size_t   size = 0;
	// allocator bookkeeping
	RingRegionBase* _region = nullptr;
	uint64_t _seq = 0;            ///< position in the region's allocation order
	uint64_t _endMark = 0;        ///< end of the region at the time of creation
	bool     _live = false;
};

struct RingRegionBase {
	uint64_t start = 0, end = 0;  ///< [oldest live allocation, end of the newest allocation)
	uint64_t frontSeq = 0, nextSeq = 0;
};

template<typename Record>
class RingSuballocator {
	static_assert(std::is_base_of<RingRecord, Record>::value, "Record must derive from RingRecord");
	struct Region : RingRegionBase { std::deque<Record> records; };
	Region _a, _b;
	Region* _r1 = &_a;            ///< region 1: grows to the end of the buffer
	Region* _r2 = &_b;            ///< region 2: grows from the start of the buffer up to region 1
	uint64_t _bufferStart = 0, _bufferEnd = 0;
	size_t _usedBytes = 0;

	Record* push(Region& r, uint64_t addr, size_t size, uint64_t endMark) {
		Record& rec = r.records.emplace_back();
		rec.deviceAddress = addr; rec.size = size; rec._region = &r; rec._seq = r.nextSeq++;
		rec._endMark = endMark; rec._live = true;
		return &rec;
	}
	void reclaim(Region& r) {
		while(!r.records.empty() && !r.records.front()._live) { r.records.pop_front(); r.frontSeq++; }
		if(!r.records.empty()) { r.start = r.records.front().deviceAddress; return; }
		if(&r == _r1) {
			if(_r2->end == _bufferStart) { r.start = r.end = _bufferStart; }
			else {  // region 2 becomes region 1; the emptied one restarts at the beginning of the buffer
				std::swap(_r1, _r2);
				_r2->start = _r2->end = _bufferStart;
			}
		}
		else r.start = r.end = _bufferStart;
	}

public:
	RingSuballocator() = default;
	RingSuballocator(uint64_t bufferStart, size_t bytes) { reset(bufferStart, bytes); }
	RingSuballocator(const RingSuballocator&) = delete;
	RingSuballocator& operator=(const RingSuballocator&) = delete;

	void reset(uint64_t bufferStart, size_t bytes) {
		_a.records.clear(); _b.records.clear();
		_a.frontSeq = _a.nextSeq = _b.frontSeq = _b.nextSeq = 0;
		_bufferStart = bufferStart; _bufferEnd = bufferStart + bytes;
		_a.start = _a.end = _b.start = _b.end = bufferStart;
		_r1 = &_a; _r2 = &_b; _usedBytes = 0;
	}

	static constexpr size_t alignmentFor(size_t numBytes) { return numBytes >= 64 ? 64 : 16; }

	/// Where would `numBytes` go?  -> {address, region 1|2} or {0, 0} when there is no contiguous room.
	std::pair<uint64_t, int> propose(size_t numBytes, size_t alignment) const {
		assert(numBytes != 0);
		const uint64_t a = alignment - 1;
		if(_r1->end + numBytes <= _bufferEnd) {
			uint64_t addr = (_r1->end + a) & ~a;
			if(addr + numBytes <= _bufferEnd) return {addr, 1};
		}
		if(_r2->end + numBytes <= _r1->start) {
			uint64_t addr = (_r2->end + a) & ~a;
			if(addr + numBytes <= _r1->start) return {addr, 2};
		}
		return {0, 0};
	}
	std::pair<uint64_t, int> propose(size_t numBytes) const { return propose(numBytes, alignmentFor(numBytes)); }

	/// Make the proposed allocation real.
	Record* commit(int region, uint64_t addr, size_t numBytes) {
		Region& r = (region == 1) ? *_r1 : *_r2;
		const uint64_t endAddr = addr + numBytes;
		Record* rec = push(r, addr, numBytes, endAddr);
		_usedBytes += endAddr - r.end;
		r.end = endAddr;
		return rec;
	}
	/// Zero-size record at the current end of a region; `addr` is recorded as its address.
	Record* pin(int region, uint64_t addr) {
		Region& r = (region == 1) ? *_r1 : *_r2;
		return push(r, addr, 0, r.end);
	}

	void release(Record* rec) {
		assert(rec && rec->_live && "record released twice");
		Region& r = *static_cast<Region*>(rec->_region);
		if(rec->size != 0) {
			const size_t idx = size_t(rec->_seq - r.frontSeq);
			const uint64_t next = (idx + 1 < r.records.size()) ? r.records[idx + 1].deviceAddress : rec->_endMark;
			_usedBytes -= size_t(next - rec->deviceAddress);
		}
		rec->_live = false;
		if(rec->_seq == r.frontSeq) reclaim(r);
	}

	int regionOf(const Record* rec) const { return rec->_region == _r1 ? 1 : 2; }
	/// Identity (0 | 1) of the storage that currently plays the role of region 1 | 2.  The roles swap when region 1 empties
	/// while region 2 holds data (reclaim), the identity of a stretch of addresses does not: state that has to follow the
	/// ADDRESSES of a region across a swap (DataMemory's pending upload runs) is keyed by this, not by the role.
	int physicalIndex(int region) const { return ((region == 1) ? _r1 : _r2) == &_a ? 0 : 1; }
	size_t usedBytes() const { return _usedBytes; }
	bool empty() const { return _a.records.empty() && _b.records.empty(); }
	uint64_t bufferStart() const { return _bufferStart; }
	uint64_t bufferEnd() const { return _bufferEnd; }
	size_t bufferSize() const { return size_t(_bufferEnd - _bufferStart); }
	uint64_t regionStart(int region) const { return region == 1 ? _r1->start : _r2->start; }
	uint64_t regionEnd(int region) const { return region == 1 ? _r1->end : _r2->end; }
};

}
