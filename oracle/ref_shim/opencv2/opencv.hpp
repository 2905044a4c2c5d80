#pragma once
namespace cv { class Mat { public: unsigned char *data = nullptr; int rows = 0, cols = 0; }; }
