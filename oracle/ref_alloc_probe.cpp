// Drives the REFERENCE'S OWN suballocator (header-only template, compiled from where it lies:
// /root/reference/src/CadR/CircularAllocationMemory.h) with a command stream read from stdin and prints the
// placements it chooses.  TEST INFRASTRUCTURE: used by oracle/make_golden.py to produce
// tests/golden/allocator_*.json, the known-answer vectors for cadr_b200/host's own allocator.
//
//   input : "<bufferBytes>\n" then one command per line:  a <id> <bytes>   |   f <id>
//   output: per command one line:  a <id> <offset | -1> <block>   |   f <id> <usedBytes>
#include <CadR/CircularAllocationMemory.h>
#include <cstdint>
#include <cstdio>
#include <map>

struct Record {                       // >= 32 bytes, first member address, second size (see the header's contract)
	uint64_t address;
	uint64_t size;
	uint64_t pad[2];
};

struct Probe : CadR::CircularAllocationMemory<Record, 200> {
	explicit Probe(uint64_t base, uint64_t bytes) {
		_bufferStartAddress = base; _bufferEndAddress = base + bytes;
		_block1StartAddress = _block1EndAddress = base;
		_block2StartAddress = _block2EndAddress = base;
	}
	Record* alloc(uint64_t bytes, int& block) {
		auto [addr, b] = allocPropose(bytes);
		block = b;
		if(b == 0) return nullptr;
		Record* r = (b == 1) ? alloc1Commit(addr, bytes) : alloc2Commit(addr, bytes);
		r->address = addr; r->size = bytes;
		return r;
	}
	void release(Record* r) { freeInternal(r); }
	uint64_t used() const { return _usedBytes; }
};

int main()
{
	const uint64_t base = 0x10000000ull;
	unsigned long long bytes;
	if(scanf("%llu", &bytes) != 1) return 1;
	Probe p(base, bytes);
	std::map<long, Record*> live;
	char c; long id; unsigned long long n;
	while(scanf(" %c %ld", &c, &id) == 2) {
		if(c == 'a') {
			if(scanf("%llu", &n) != 1) return 1;
			int block = 0;
			Record* r = p.alloc(n, block);
			if(r) { live[id] = r; printf("a %ld %lld %d\n", id, (long long)(r->address - base), block); }
			else printf("a %ld -1 0\n", id);
		}
		else if(c == 'f') {
			auto it = live.find(id);
			if(it == live.end()) { printf("f %ld -1\n", id); continue; }   // the allocation had failed: nothing to free
			p.release(it->second);
			live.erase(it);
			printf("f %ld %llu\n", id, (unsigned long long)p.used());
		}
	}
	return 0;
}
