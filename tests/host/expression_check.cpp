// Host-only check of the prescription expression parser: evaluates every expression of argv[1] (one per line) at the
// times given as further arguments and prints the values.
#include <DEM/utils/Expression.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

int main(int argc, char** argv) {
    std::ifstream f(argv[1]);
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        try {
            deme::TimeExpression e(line);
            printf("%d", (int)e.IsConstant());
            for (int k = 2; k < argc; k++) printf(" %.17g", e.Eval(atof(argv[k])));
            printf("\n");
        } catch (const std::exception& ex) {
            printf("ERROR\n");
        }
    }
    return 0;
}
