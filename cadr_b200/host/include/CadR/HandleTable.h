// HandleTable — 64-bit handles (from 1, never recycled) resolved on the device through 1-3 levels of 2048-entry
// tables.  Reference: src/CadR/HandleTable.{h,cpp}.  Table nodes live in DataStorage like any other allocation and
// follow realloc-on-write: the first write to a node in a frame moves it, and the move is propagated to its parent
// up to the root (HandleTable.cpp:58-68,348-378), whose address the processing kernel receives every frame.
#pragma once
#include <CadR/DataAllocation.h>
#include <array>
#include <memory>

namespace CadR {

class HandleTable {
public:
	static constexpr unsigned numHandlesPerTable = 2048;
	static constexpr unsigned handleBitsLevelShift = 11;
	static constexpr unsigned handleBitsLevelMask = 0x07ff;
private:
	struct Node {
		HandlelessAllocation allocation;
		std::array<uint64_t, numHandlesPerTable> entries{};
		std::array<std::unique_ptr<Node>, numHandlesPerTable>* children = nullptr;  // routing nodes only
		Node* parent = nullptr;
		unsigned indexInParent = 0;
		explicit Node(DataStorage& storage, bool routing);
		~Node();
		void init();                                   // zero-filled 16 KiB block (HandleTable.cpp:40-55)
	};
	DataStorage* _storage;
	std::unique_ptr<Node> _root;
	uint64_t _highestHandle = 0;
	unsigned _handleLevel = 0;
	void setEntry(Node& node, unsigned index, uint64_t value);   // with relocation cascade to the root
	Node* leafFor(uint64_t handle) const;
	std::unique_ptr<Node> makeNode(bool routing);
public:
	explicit HandleTable(DataStorage& storage) noexcept : _storage(&storage) {}
	~HandleTable() noexcept { destroyAll(); }
	uint64_t create();
	uint64_t create(uint64_t deviceAddress) { uint64_t h = create(); set(h, deviceAddress); return h; }
	void destroy(uint64_t) noexcept {}                            // handles are never recycled (HandleTable.h:101)
	void destroyAll() noexcept;
	void set(uint64_t handle, uint64_t addr);
	unsigned handleLevel() const { return _handleLevel; }
	uint64_t rootTableDeviceAddress() const;
	uint64_t highestHandle() const { return _highestHandle; }
};

}
