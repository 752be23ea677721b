// Minimal XML DOM for VXA / VXD / report files (host side, parse-time only).
// Replaces the reference's use of tinyxml (Voxelyze3/src/old/XML_Rip.cpp) and boost::property_tree
// (VX3_SimulationManager.cu:277-293) for the subset those files use: elements, attributes, text,
// CDATA, comments, the XML declaration and the five predefined entities.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace vx3b {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text; // concatenated character data + CDATA directly under this element
    std::vector<std::unique_ptr<XmlNode>> children;

    const XmlNode *child(const std::string &n) const {
        for (auto &c : children)
            if (c->name == n)
                return c.get();
        return nullptr;
    }
    XmlNode *child(const std::string &n) {
        for (auto &c : children)
            if (c->name == n)
                return c.get();
        return nullptr;
    }
    // dotted path, first match at every level ("Simulator.Integration.DtFrac")
    const XmlNode *find(const std::string &path) const {
        const XmlNode *cur = this;
        size_t p = 0;
        while (cur && p <= path.size()) {
            size_t q = path.find('.', p);
            std::string part = path.substr(p, q == std::string::npos ? std::string::npos : q - p);
            if (!part.empty())
                cur = cur->child(part);
            if (q == std::string::npos)
                break;
            p = q + 1;
        }
        return cur;
    }
    const std::string *attr(const std::string &n) const {
        for (auto &a : attrs)
            if (a.first == n)
                return &a.second;
        return nullptr;
    }
    std::unique_ptr<XmlNode> clone() const {
        auto r = std::make_unique<XmlNode>();
        r->name = name;
        r->attrs = attrs;
        r->text = text;
        for (auto &c : children)
            r->children.push_back(c->clone());
        return r;
    }
};

std::string xml_trim(const std::string &s);
// Parses a document; returns a synthetic root whose children are the top-level elements.
std::unique_ptr<XmlNode> xml_parse(const std::string &text, std::string *err);
// Serialises `node` (an element) with 2-space indentation.
std::string xml_write(const XmlNode &node, bool declaration = true);
// boost::property_tree::put_child semantics on a dotted path: replace the first child with that
// name at each level, creating missing levels (Utils/ctool.h:49-57).
void xml_put_child(XmlNode &root, const std::string &path, std::unique_ptr<XmlNode> subtree);

} // namespace vx3b
