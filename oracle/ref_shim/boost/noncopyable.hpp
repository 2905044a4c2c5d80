#pragma once
// stand-in: boost::noncopyable as the reference's Frame / VisualPoint use it (include/frame.h:16, include/visual_point.h:16)
namespace boost { class noncopyable { protected: noncopyable() = default; ~noncopyable() = default; noncopyable(const noncopyable &) = delete; noncopyable &operator=(const noncopyable &) = delete; }; }
