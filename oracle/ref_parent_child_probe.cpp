// Drives the REFERENCE'S OWN parent/child relation lists (header-only templates over boost::intrusive, compiled from
// where they lie: /root/reference/src/CadR/ParentChildList.h, declared exactly as CadR::StateSet declares them,
// StateSet.h:77-79) with a command stream and prints the resulting child and parent orders.  TEST INFRASTRUCTURE:
// oracle/make_golden.py turns its output into tests/golden/parent_child_kat.json.gz, the known-answer vectors for the
// facade's StateSet::childList / parentList - whose iteration order is the order in which StateSets are flattened.
//
//   input : "<nodes>\n" then one command per line
//             ac <p> <c>   p.childList.append(c)        ap <c> <p>   c.parentList.append(p)
//             rc <p> <k>   remove p's k-th child link   rp <c> <k>   remove c's k-th parent link
//             cc <p>       p.childList.clear()          cp <c>       c.parentList.clear()
//             s            print the state
//   output: per "s" one line: for every node "<id>:c=<ids,>;p=<ids,>" separated by spaces
#include <CadR/ParentChildList.h>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

struct Node {
	int id;
	static const CadR::ParentChildListOffsets offsets;
	CadR::ChildList<Node, offsets> childList;
	CadR::ParentList<Node, offsets> parentList;
};
const CadR::ParentChildListOffsets Node::offsets{ offsetof(Node, parentList), offsetof(Node, childList) };

int main()
{
	int n = 0;
	if(scanf("%d", &n) != 1 || n <= 0) return 2;
	std::vector<std::unique_ptr<Node>> nodes;
	for(int i = 0; i < n; i++) { nodes.push_back(std::make_unique<Node>()); nodes.back()->id = i; }
	char cmd[8];
	while(scanf("%7s", cmd) == 1) {
		int a = 0, b = 0;
		if(!strcmp(cmd, "s")) {
			for(auto& nd : nodes) {
				printf("%d:c=", nd->id);
				for(auto it = nd->childList.begin(); it != nd->childList.end(); ++it) printf("%d,", (*it).id);
				printf(";p=");
				for(auto it = nd->parentList.begin(); it != nd->parentList.end(); ++it) printf("%d,", (*it).id);
				printf(" ");
			}
			printf("\n");
			continue;
		}
		if(cmd[0] == 'c') { if(scanf("%d", &a) != 1) return 2; }
		else if(scanf("%d %d", &a, &b) != 2) return 2;
		Node& x = *nodes[size_t(a)];
		if(!strcmp(cmd, "ac")) x.childList.append(*nodes[size_t(b)]);
		else if(!strcmp(cmd, "ap")) x.parentList.append(*nodes[size_t(b)]);
		else if(!strcmp(cmd, "rc")) { auto it = x.childList.begin(); for(int k = 0; k < b; k++) ++it; x.childList.remove(it); }
		else if(!strcmp(cmd, "rp")) { auto it = x.parentList.begin(); for(int k = 0; k < b; k++) ++it; x.parentList.remove(it); }
		else if(!strcmp(cmd, "cc")) x.childList.clear();
		else if(!strcmp(cmd, "cp")) x.parentList.clear();
		else return 2;
	}
	return 0;
}
