// Generator of tests/golden/libtorch_model.ot (+ the raw blob hashed in libtorch_model.json).  Build (this image):
//   T=$(python -c 'import torch,os;print(os.path.dirname(torch.__file__))')
//   g++ -std=c++17 -O1 -I$T/include -I$T/include/torch/csrc/api/include make_ot_fixture.cpp -o mk -L$T/lib -ltorch -ltorch_cpu -lc10 -Wl,-rpath,$T/lib
//   ./mk libtorch_model.ot blob.bin
// What tch 0.4.1's VarStore::save does (tch/src/nn/var_store.rs -> Tensor::save_multi -> torch-sys at_save_multi):
// torch::serialize::OutputArchive; archive.write(name, tensor) per named variable; archive.save_to(path).
#include <torch/serialize/archive.h>
#include <torch/torch.h>
#include <cstdio>
int main(int argc, char** argv) {
    torch::manual_seed(7);
    const char* names[] = {"l_1.weight","l_1.bias","l_2.weight","l_2.bias","l_3.weight","l_3.bias","l_4.weight","l_4.bias","l_5.weight","l_5.bias"};
    int dims[][2] = {{128,63},{128,0},{96,128},{96,0},{64,96},{64,0},{48,64},{48,0},{12,48},{12,0}};
    torch::serialize::OutputArchive ar;
    FILE* f = fopen(argv[2], "wb");
    for (int i = 0; i < 10; ++i) {
        auto t = dims[i][1] ? torch::randn({dims[i][0], dims[i][1]}) : torch::randn({dims[i][0]});
        ar.write(names[i], t);
        fwrite(t.data_ptr<float>(), 4, t.numel(), f);
    }
    fclose(f);
    ar.save_to(argv[1]);
    return 0;
}
