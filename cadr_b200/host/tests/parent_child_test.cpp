// The scenario of the reference's tests/ParentChildTest.cpp (:12-38) against the facade's StateSet lists, plus the
// two-way bookkeeping the reference gets from its intrusive relation objects.
#include <CadR/CadR.h>
#include <cstdio>
#include <cstdlib>

using namespace CadR;
#define REQUIRE(c) do { if(!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while(0)

int main()
{
	Renderer r(Renderer::addressSpaceOnly);
	StateSet ss1(r), ss2(r), ss3(r);

	REQUIRE(ss1.childList.size() == 0 && ss1.childList.empty());
	auto it1 = ss1.childList.append(ss2);
	REQUIRE(ss1.childList.size() == 1 && ss2.parentList.size() == 1 && &ss2.parentList.front() == &ss1);
	ss1.childList.remove(it1);
	REQUIRE(ss1.childList.empty() && ss2.parentList.empty());
	ss1.childList.clear();
	for(const StateSet& child : ss1.childList) child.renderer();

	auto it2 = ss2.parentList.append(ss1);
	REQUIRE(ss2.parentList.size() == 1 && ss1.childList.size() == 1 && &ss1.childList.back() == &ss2);
	ss2.parentList.remove(it2);
	REQUIRE(ss2.parentList.empty() && ss1.childList.empty());
	ss2.parentList.clear();
	for(const StateSet& parent : ss2.parentList) parent.renderer();

	// a StateSet with two parents is visited once per parent
	ss1.childList.append(ss3);
	ss2.childList.append(ss3);
	REQUIRE(ss3.parentList.size() == 2);
	ss1.childList.clear();
	REQUIRE(ss3.parentList.size() == 1 && &ss3.parentList.front() == &ss2);
	printf("parent_child_test ok\n");
	return 0;
}
