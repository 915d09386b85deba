// examples/main.cpp - the reference's driver (main.cpp:11-28) written against the facade headers in include/PathTracerAP/.
// The reference's own main.cpp compiles against the same headers unchanged (tests/test_facade.py does exactly that where
// /root/reference exists); this copy exists so that the GPU box, which has no reference tree, can build and run the flow.
//   g++ -std=c++17 -Iinclude/PathTracerAP examples/main.cpp -Lpathtracerap_b200 -lptap -Wl,-rpath,$PWD/pathtracerap_b200 -o pt_main
//   ./pt_main [Config.txt]          (without an argument: the hard-coded scene from "Input data/" in the current directory)
#include "Renderer.h"
#include "Scene.h"

int main(int argc, char** argv)
{
    try {
        Scene* scene = new Scene(argc > 1 ? argv[1] : "Input data\\lucy.obj");
        Renderer* renderer = new Renderer();
        renderer->allocateOnGPU(*scene);
        delete scene;                               // the renderer holds its own copy, as in the reference
        renderer->renderLoop();
        renderer->renderImage();
        renderer->free();
        delete renderer;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
