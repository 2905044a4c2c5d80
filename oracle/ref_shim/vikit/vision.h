#pragma once
// Stand-ins for vikit_common/vision.h: the three functions src/vio.cpp calls, restated from the published code.
#include <cmath>
#include <opencv2/opencv.hpp>
namespace vk {
// bilinear sample of an 8-bit image, weights from the fractional part of (u, v)
inline float interpolateMat_8u(const cv::Mat &mat, float u, float v) {
  const int x = (int)std::floor(u), y = (int)std::floor(v);
  const float subpix_x = u - x, subpix_y = v - y;
  const float w00 = (1.0f - subpix_x) * (1.0f - subpix_y), w01 = (1.0f - subpix_x) * subpix_y, w10 = subpix_x * (1.0f - subpix_y), w11 = 1.0f - w00 - w01 - w10;
  const int stride = (int)mat.step.p[0];
  const unsigned char *ptr = mat.data + y * stride + x;
  return w00 * ptr[0] + w01 * ptr[stride] + w10 * ptr[1] + w11 * ptr[stride + 1];
}
inline float shiTomasiScore(const cv::Mat &img, int u, int v) {
  float dXX = 0.0, dYY = 0.0, dXY = 0.0;
  const int halfbox_size = 4, box_size = 2 * halfbox_size, box_area = box_size * box_size;
  const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
  if (x_min < 1 || x_max >= img.cols - 1 || y_min < 1 || y_max >= img.rows - 1) return 0.0;
  const int stride = (int)img.step.p[0];
  for (int y = y_min; y < y_max; ++y) {
    const unsigned char *ptr_left = img.data + stride * y + x_min - 1, *ptr_right = img.data + stride * y + x_min + 1;
    const unsigned char *ptr_top = img.data + stride * (y - 1) + x_min, *ptr_bottom = img.data + stride * (y + 1) + x_min;
    for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
      const float dx = *ptr_right - *ptr_left, dy = *ptr_bottom - *ptr_top;
      dXX += dx * dx, dYY += dy * dy, dXY += dx * dy;
    }
  }
  dXX = dXX / (2.0 * box_area), dYY = dYY / (2.0 * box_area), dXY = dXY / (2.0 * box_area);
  return 0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}
inline void halfSample(const cv::Mat &in, cv::Mat &out) {
  for (int y = 0; y < out.rows; y++)
    for (int x = 0; x < out.cols; x++) {
      const unsigned char *p = in.data + (2 * y) * in.step.p[0] + 2 * x;
      out.data[y * out.step.p[0] + x] = (unsigned char)((p[0] + p[1] + p[in.step.p[0]] + p[in.step.p[0] + 1]) / 4);
    }
}
}  // namespace vk
