// tests/cpp/test_bcqp.cpp -- BCQPSolver's internal self test through the C++ mirror (include/alens_b200/BCQPSolver.hpp),
// driven like SimToolbox/Constraint/BCQPSolver_test.cpp:19-36:  test_bcqp <localSize> <diagonal> <seed> <solverChoice> <out.bin>
// out.bin: the dense matrix A (n*n doubles, row major), b, lb, ub, the solution (n doubles each)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "alens_b200/BCQPSolver.hpp"

int main(int argc, char **argv) {
    if (argc < 6) return 1;
    const int n = atoi(argv[1]);
    const double diagonal = atof(argv[2]);
    const unsigned seed = (unsigned)atoi(argv[3]);
    const int choice = atoi(argv[4]);
    alens_ctx *ctx = nullptr;
    if (alens_create(0, 0, 1, &ctx) != ALENS_OK) {
        fprintf(stderr, "%s\n", alens_last_error(nullptr));
        return 2;
    }
    try {
        BCQPSolver test(n, diagonal, ctx, seed);
        Teuchos::RCP<TV> x;
        test.selfTest(1e-7, 3000, choice, &x);
        const auto *A = dynamic_cast<const TCMAT *>(test.getOperator().get());
        std::vector<double> dense((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++)
            for (long long k = A->rowPtr()[i]; k < A->rowPtr()[i + 1]; k++) dense[(size_t)i * n + A->colInd()[k]] = A->values()[k];
        FILE *f = fopen(argv[5], "wb");
        fwrite(dense.data(), 8, dense.size(), f);
        fwrite(test.getB()->data(), 8, n, f);
        fwrite(test.getLowerBound()->data(), 8, n, f);
        fwrite(test.getUpperBound()->data(), 8, n, f);
        fwrite(x->data(), 8, n, f);
        fclose(f);
    } catch (const std::exception &e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 3;
    }
    alens_destroy(ctx);
    return 0;
}
