// The facade's StateSet::childList / parentList driven by the command stream of oracle/ref_parent_child_probe.cpp
// (same protocol, same output format), so that tests/test_host_cpu.py can compare it line by line with what the
// reference's own ParentChildList.h produced (tests/golden/parent_child_kat.json.gz).
#include <CadR/CadR.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

using namespace CadR;

int main()
{
	int n = 0;
	if(scanf("%d", &n) != 1 || n <= 0) return 2;
	Renderer r(Renderer::addressSpaceOnly);
	std::vector<std::unique_ptr<StateSet>> nodes;
	std::map<const StateSet*, int> idOf;
	for(int i = 0; i < n; i++) { nodes.push_back(std::make_unique<StateSet>(r)); idOf[nodes.back().get()] = i; }
	char cmd[8];
	while(scanf("%7s", cmd) == 1) {
		int a = 0, b = 0;
		if(!strcmp(cmd, "s")) {
			for(auto& nd : nodes) {
				printf("%d:c=", idOf[nd.get()]);
				for(StateSet& c : nd->childList) printf("%d,", idOf[&c]);
				printf(";p=");
				for(StateSet& p : nd->parentList) printf("%d,", idOf[&p]);
				printf(" ");
			}
			printf("\n");
			continue;
		}
		if(cmd[0] == 'c') { if(scanf("%d", &a) != 1) return 2; }
		else if(scanf("%d %d", &a, &b) != 2) return 2;
		StateSet& x = *nodes[size_t(a)];
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
