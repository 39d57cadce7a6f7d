// Reader for the SLR scene-description language: hand-written lexer + precedence-climbing parser +
// tree-walking interpreter (the reference uses flex/bison generated code, SceneLexer.l /
// SceneParser.yy; neither tool exists in this image and nothing here is generated).
#pragma once
#include "../renderer.h"
#include "../scene.h"
#include <string>

namespace slr {

// libSLRSceneGraph/API.hpp:20 -- parses and executes `filePath`, filling the scene graph and the
// rendering context (renderer choice + settings). Asset paths are relative to the scene file.
// Returns false and sets *error on a syntax or execution error.
bool readScene(const std::string& filePath, Scene* scene, RenderingContext* context, std::string* error, bool rgbMode = false);

}  // namespace slr
